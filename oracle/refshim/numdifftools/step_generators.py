class MaxStepGenerator:
    """Referenced in a default argument at gwfast/signal.py:1102; never called."""

    def __init__(self, *a, **k):
        pass
