#!/usr/bin/env python
"""Benchmark of the Fisher/SNR hot path (BASELINE.json metric: Fisher events/s, IMRPhenomD, ET+2CE).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, all host threads)

A "step" = one pass of the hot path over one batch of the synthetic catalog: DetNet.SNR + DetNet.FisherMatr
(res=1000, spacing='geom', use_chi1chi2=True, Earth rotation on), BASELINE.json configs[1]: 10^4 IMRPhenomD BBH events on
ET (triangle) + 2 CE per GPU (weak scaling).  Rank 0 prints ONE JSON line.

  value     kernel-path events/s, event parameters already resident in HBM (CUDA events, max over ranks; for N>1 the
            final NCCL all-gather of the packed Fisher matrices is inside the timed region)
  e2e       the same metric through the public API with HOST numpy arrays in and out (H2D/D2H inside the timed region); for N>1
            every rank calls the API on its shard and the all-gather is taken from the engine's device-resident Fisher matrices
            (gwfast_b200.parallel.fisher_with_device_gather): hosts read their own shard, the full matrix stays in HBM
  roofline  FP64 (the path is FP64-FMA/transcendental bound): algorithmic FLOP/event x events / duration of
            the dominant kernel (fisher_kernel, timed alone via GWF_OPT_REUSE_WORKSPACE) against the DFMA peak measured
            in the same run (gwf_fp64_peak) -- nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s is also reported
  cpu_baseline  the oracle port (numpy + forward-mode duals, a restatement of the reference's own CPU algorithm) on a
            bounded sample of the same catalog, all host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = 'IMRPhenomD BBH, ET(triangle)+CE1Id+CE2NM, 10^4 events/GPU, 11-param Fisher + SNR, Earth rotation on, res=1000'
EVENTS_PER_GPU = 10000
RES = 1000
# algorithmic FLOP per event for this configuration (frozen in DESIGN.md): add/mul = 1, FMA = 2,
# div/sqrt/transcendental = 1; 4 evaluated arms, 5 Grams, nP = 11, res = 1000
FLOP_PER_EVENT = 2.676e6
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
# DRAM traffic of one fisher_kernel launch of this workload, bytes (ncu --set full, profiles/r01i_kernels_ncu.md): 34.71 MB read + 0.57 MB written
DRAM_BYTES_PER_LAUNCH = 35.27e6


# ----------------------------------------------------------------------------------------------- CPU arm
def _cpu_chunk(args):
    lo, hi = args
    import warnings
    warnings.filterwarnings('ignore')
    from gwfast_b200 import synthetic
    from oracle.port import waveforms as PW, detector as PD
    ev = {k: v[lo:hi] for k, v in synthetic.bbh_catalog(EVENTS_PER_GPU, synthetic.SEEDS['C2']).items()}
    net = PD.Network(synthetic.build_network(PD.Detector, PW.IMRPhenomD(), 'ET+2CE', useEarthMotion=True, fmin=2.))
    snr = net.SNR(dict(ev), res=RES)
    F = net.FisherMatr(dict(ev), res=RES)
    return float(snr.sum() + F[0, 0].sum())


class CpuPool:
    """worker pool running the oracle port; one process per host core (mirrors the reference's --npools batches,
    run/calculate_forecasts_from_catalog.py:152-186, 1016-1024)."""

    def __init__(self, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context('spawn').Pool(cores)
        self.pool.map(_cpu_chunk, [(0, 2)] * cores)        # warm the workers (imports, table loads)

    def run(self, n_events, offset=0):
        """oracle port on `n_events` events of the catalog split in contiguous batches; returns (events, seconds)."""
        per = max(1, n_events // self.cores)
        chunks = [(offset + i * per, offset + (i + 1) * per) for i in range(self.cores)]
        t = time.perf_counter()
        self.pool.map(_cpu_chunk, chunks)
        return per * self.cores, time.perf_counter() - t

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    cores = os.cpu_count() or 1
    per_step = 32 * cores                                  # bounded sample: 32 events per core per step (~1.3 s)
    pool = CpuPool(cores)
    n_tot = t_tot = 0.
    for i in range(args.warmup + args.steps):
        n, dt = pool.run(per_step, offset=(i * per_step) % (EVENTS_PER_GPU - per_step))
        if i >= args.warmup:
            n_tot += n
            t_tot += dt
    pool.close()
    value = n_tot / t_tot
    sample = '%d events/step x %d steps of the C2 catalog, multiprocessing.Pool(%d), numpy+dual oracle port' % (per_step, args.steps, cores)
    line = dict(impl='reference', metric='fisher_events_per_s', value=value, unit='events/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_tot / max(1, args.steps), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=WORKLOAD, note='reference CPU path = oracle port (the Python/JAX reference cannot travel to the GPU box)'),
                cpu_baseline=dict(value=value, unit='events/s', cores=cores, kind='port', sample=sample),
                e2e=dict(value=value, unit='events/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  Read in-process through NVML
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints): an nvidia-smi child
    polling every 100 ms stalls the CUDA driver for milliseconds at a time, which the host-timed e2e region then pays for.
    Falls back to the nvidia-smi loop if NVML cannot be loaded."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
    BITS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def _run_nvml(self):
        import pynvml as N
        N.nvmlInit()
        # LOCAL_RANK indexes the visible devices; NVML enumerates all of them
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        idx = int(vis.split(',')[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(',')) else self.index
        h = N.nvmlDeviceGetHandleByIndex(idx)
        mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
        reasons = getattr(N, 'nvmlDeviceGetCurrentClocksEventReasons', None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            sm = N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)
            mask = int(reasons(h))
            self.rows.append([str(sm), str(mx), ''] + ['Active' if mask & bit else 'Not Active' for _, bit in self.BITS])
            time.sleep(0.02)

    def _run_smi(self):
        try:
            p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        while not self.stop_flag:
            ln = p.stdout.readline()
            if not ln:
                break
            self.rows.append([x.strip() for x in ln.split(',')])
        p.terminate()

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            self._run_smi()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = [n for n, _ in self.BITS]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == 'Active' for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def run_engine(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from gwfast_b200 import waveforms, signal, network, synthetic, parallel, _engine, _capi as K
    import ctypes as C

    # weak scaling: every rank owns its own 10^4-event shard of a world*10^4-event catalog
    full = synthetic.bbh_catalog(EVENTS_PER_GPU * world, synthetic.SEEDS['C2'] if world == 1 else synthetic.SEEDS['C5'])
    ev = {k: np.ascontiguousarray(v[rank * EVENTS_PER_GPU:(rank + 1) * EVENTS_PER_GPU]) for k, v in full.items()}
    n = EVENTS_PER_GPU
    wf = waveforms.IMRPhenomD()
    sigs = synthetic.build_network(signal.GWSignal, wf, 'ET+2CE', useEarthMotion=True, fmin=2.)
    net = network.DetNet(sigs, verbose=False)
    st = _engine.state()
    lib = st.lib
    dev = st.device
    stream = torch.cuda.current_stream(dev)
    sp = C.c_void_p(stream.cuda_stream)

    # ---- resident inputs and reusable outputs for the kernel-path measurement
    model = wf._descriptor(ev)
    dets = [s._detector_struct(i) for i, s in enumerate(sigs.values())]
    handles = [s._psd_handle() for s in sigs.values()]
    darr, parr = _engine._call_arrays(dets, handles)
    dev_ev, host_ev, evs, _ = _engine._upload(st, signal._engine_events(wf, ev), n, K.EVENT_KEYS)
    nP, npack, narms = 11, 66, 5
    ws = _engine._workspace(st, lib.gwf_workspace_bytes(C.byref(model), n))
    packed = torch.empty((n, npack), dtype=torch.float64, device=dev)
    snr2 = torch.empty((n,), dtype=torch.float64, device=dev)
    snr2_arm = torch.empty((narms, n), dtype=torch.float64, device=dev)
    fullF = torch.empty((nP, nP, n), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, n, npack), dtype=torch.float64, device=dev) if world > 1 else None
    opts = K.gwf_opts(RES, 0, 0, 0)
    opts_reuse = K.gwf_opts(RES, K.GWF_OPT_REUSE_WORKSPACE, 0, 0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def kernel_step():
        K.check(lib.gwf_snr(C.byref(model), darr, len(dets), parr, len(handles), C.byref(evs), n, C.byref(opts), C.c_void_p(snr2_arm.data_ptr()),
                            C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_snr')
        K.check(lib.gwf_fisher(C.byref(model), darr, len(dets), parr, len(handles), C.byref(evs), n, C.byref(opts), C.c_void_p(packed.data_ptr()),
                               C.c_void_p(snr2.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_fisher')
        K.check(lib.gwf_unpack_fisher(C.c_void_p(packed.data_ptr()), n, nP, C.c_void_p(fullF.data_ptr()), sp), 'gwf_unpack_fisher')
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1))
        return 5                                                                    # prologue x2, snr, fisher, unpack

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()

    # ---- kernel path (value)
    for _ in range(max(3, args.warmup)):
        kernel_step()
    barrier()
    t_dev = 0.0
    launches = 0
    for _ in range(args.steps):
        flush.fill_(1)                                                              # flush L2 between timed iterations
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        launches += kernel_step()
        e1.record(stream)
        barrier()
        t_dev += e0.elapsed_time(e1) * 1e-3
    # ---- dominant kernel alone (fisher_kernel): records are still in the workspace
    t_main = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        K.check(lib.gwf_fisher(C.byref(model), darr, len(dets), parr, len(handles), C.byref(evs), n, C.byref(opts_reuse), C.c_void_p(packed.data_ptr()),
                               C.c_void_p(snr2.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), sp), 'gwf_fisher')
        e1.record(stream)
        torch.cuda.synchronize()
        t_main += e0.elapsed_time(e1) * 1e-3
    # ---- end to end through the public API, host numpy in / host numpy out
    for _ in range(max(3, args.warmup)):
        # results are held like in the timed loop, so the pinned-buffer pool reaches its steady state here (the previous
        # step's arrays are still alive when the next call allocates: one extra cudaHostAlloc, ~9 ms, the first time)
        s_ = net.SNR(dict(ev), res=RES)
        if world > 1:
            F_, g = parallel.fisher_with_device_gather(net, dict(ev), n * world, dist, res=RES)
        else:
            F_ = net.FisherMatr(dict(ev), res=RES)
    barrier()
    t_e2e = 0.0
    h2d = d2h = 0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t = time.perf_counter()
        s_ = net.SNR(dict(ev), res=RES)
        if world > 1:
            # host arrays in, this rank's host arrays out, plus the final gather of the results (north star: one NCCL all-gather
            # of Fisher matrices) taken from the engine's device-resident result: the full matrix stays in HBM on every rank
            F_, g = parallel.fisher_with_device_gather(net, dict(ev), n * world, dist, res=RES)
        else:
            F_ = net.FisherMatr(dict(ev), res=RES)
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t
        if os.environ.get('GWF_BENCH_DEBUG'):
            print('e2e step %.3f ms' % (1e3 * (time.perf_counter() - t)), file=sys.stderr)
        h2d = 2 * 13 * n * 8
        d2h = (F_.size + n + narms * n) * 8
    sampler.stop_flag = True
    time.sleep(0.15)

    # ---- measured FP64 peak, same run
    peak = C.c_double(0.)
    K.check(lib.gwf_fp64_peak(50.0, C.byref(peak), sp), 'gwf_fp64_peak')

    # max over ranks
    tt = torch.tensor([t_dev, t_main, t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev, t_main, t_e2e = [float(x) for x in tt.tolist()]
    total_events = n * world * args.steps
    value = total_events / t_dev
    e2e_value = total_events / t_e2e
    achieved = FLOP_PER_EVENT * n * args.steps / t_main / 1e12
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)
        ns, dt = pool.run(256 * cores)                      # bounded sample: ~10 s of CPU work on all host cores
        pool.close()
        cpu = dict(value=ns / dt, unit='events/s', cores=cores, kind='port',
                   sample='%d events of the C2 catalog (%.1f s wall), multiprocessing.Pool(%d), numpy+dual oracle port' % (ns, dt, cores))
    if rank == 0:
        line = dict(metric='fisher_events_per_s', value=value, unit='events/s', n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                    config=dict(workload=WORKLOAD, events_per_gpu=n, res=RES, l2='flushed between timed iterations (256 MiB fill)',
                                parallelism='events sharded contiguously, one rank per GPU' + (', final NCCL all-gather of packed Fisher' if world > 1 else '')),
                    e2e=dict(value=e2e_value, unit='events/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches,
                    roofline=dict(bound='fp64', kernel='fisher_kernel<IMRPhenomD,NT=4>', achieved=achieved, peak=float(peak.value), unit='TFLOP/s',
                                  frac=achieved / float(peak.value) if peak.value > 0 else None, traffic=DRAM_BYTES_PER_LAUNCH,
                                  traffic_source='dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture profiles/r01i (records + EventAux + PSD windows + events in, packed Fisher out)',
                                  peak_source='measured in this run (gwf_fp64_peak DFMA chain); MEASURED_PEAKS.json has no FP64 entry',
                                  nominal_peak=FP64_NOMINAL_TFLOPS, frac_of_nominal=achieved / FP64_NOMINAL_TFLOPS,
                                  flop_per_event=FLOP_PER_EVENT, kernel_ms=1e3 * t_main / args.steps),
                    clocks=sampler.summary())
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
