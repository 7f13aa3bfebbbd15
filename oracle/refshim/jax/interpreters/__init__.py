"""jax.interpreters stub (the reference imports `xla` and never uses it)."""
xla = None
