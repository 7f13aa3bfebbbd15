#!/usr/bin/env python
"""Benchmark of the Fisher/SNR hot path (BASELINE.json metric: Fisher events/s, IMRPhenomD, ET+2CE).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port, all host threads)

A "step" = one pass of the hot path over one batch of the synthetic catalog: the network SNR and the full Fisher matrix of every
event (res=1000, spacing='geom', use_chi1chi2=True, Earth rotation on), BASELINE.json configs[1]: 10^4 IMRPhenomD BBH events on
ET (triangle) + 2 CE per GPU (weak scaling).  Both results come from ONE fused launch (gwf_fisher_ex with snr2_integ /
DetNet.FisherMatr(return_SNR=True)): the Fisher kernel integrates |h|^2/S_n anyway.  Rank 0 prints ONE JSON line.

  value     kernel-path events/s, event parameters already resident in HBM (CUDA events, max over ranks).  For N>1 the final gather
            of the packed Fisher matrices is inside the timed region: the Fisher kernel stores every finished row straight into the
            peers' gathered buffers over NVLink (gwf_fisher_out.peer_fisher, peer memory mapped with CUDA IPC), so the gather overlaps
            the computation; if the peers cannot be mapped it falls back to dist.all_gather_into_tensor and says so in config.gather
  e2e       the same metric through the public API with HOST numpy arrays in and out (H2D/D2H inside the timed region); for N>1
            every rank calls the API on its shard and the gather is done by the engine's own Fisher kernels (peer stores over NVLink)
  roofline  FP64 (the path is FP64-FMA/transcendental bound): algorithmic FLOP/event x events / duration of
            the dominant kernel (fisher_kernel, timed alone via GWF_OPT_REUSE_WORKSPACE) against the DFMA peak measured
            in the same run (gwf_fp64_peak) -- nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2 TFLOP/s is also reported
  other_configs  the other BASELINE.json configurations, same measurement (kernel path, dominant kernel alone, end to end): C1 TaylorF2 /
            one L, C3 NRTidalv2 / ET+2CE, C4 IMRPhenomHM / LVK-O4 (10^5 events split over the N ranks), C5 = 10^6 IMRPhenomD events on
            ET+2CE split over the N ranks (the north-star target run; IMRPhenomXAS does not exist in the reference)
  separate_calls  the step as two API calls (gwf_snr + gwf_fisher / DetNet.SNR + DetNet.FisherMatr), for comparison with round 1
  multi_gpu_check  N>1: every rank recomputes 64 events of another rank's shard and compares them bitwise with the gathered result
  cpu_baseline  the oracle port (numpy + forward-mode duals, a restatement of the reference's own CPU algorithm) on a
            bounded sample of the same catalog, all host cores, median of 3
"""
import argparse
import ctypes as C
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
# algorithmic FLOP per event (frozen in BASELINE.md 4 / DESIGN.md 4): add/mul = 1, FMA = 2, div/sqrt/transcendental = 1
FLOP_PER_EVENT = 2.676e6
FLOP_MODEL = {'C1': 0.674e6, 'C2': 2.676e6, 'C3': 3.444e6, 'C4': 4.176e6, 'C5': 2.676e6}
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
# DRAM traffic of one fisher_kernel launch of this workload, bytes (ncu --set full, profiles/r02d_fisher_d_ncu_raw.csv: 34.70 MB read + 0.32 MB
# written; the round's last capture, profiles/r02k_fisher_d_ncu_raw.csv, ran without bench.py's L2 flush and read 29.35 MB)
DRAM_BYTES_PER_LAUNCH = 35.02e6


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
    per_step = 64 * cores                                  # bounded sample: 64 events per core per step (~3 s)
    pool = CpuPool(cores)
    rates, n_tot, t_tot = [], 0., 0.
    for i in range(args.warmup + args.steps):
        n, dt = pool.run(per_step, offset=(i * per_step) % (EVENTS_PER_GPU - per_step))
        if i >= args.warmup:
            rates.append(n / dt)
            n_tot += n
            t_tot += dt
    pool.close()
    value = float(np.median(rates))                        # per-step rates: the median is robust against a noisy neighbour on the host
    sample = '%d events/step x %d steps of the C2 catalog (median of the per-step rates), multiprocessing.Pool(%d), numpy+dual oracle port' % (
        per_step, args.steps, cores)
    line = dict(impl='reference', metric='fisher_events_per_s', value=value, unit='events/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_tot / max(1, args.steps), higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload=WORKLOAD, events_per_gpu=EVENTS_PER_GPU, res=RES,
                            note='reference CPU path = oracle port (the Python/JAX reference cannot travel to the GPU box); bounded sample of the same catalog'),
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


class Case:
    """One configuration with its inputs resident in HBM: the C-ABI calls of a step on the current stream, and the API objects."""

    def __init__(self, tag, model_name, netname, ev, rot=True, fmin=2.):
        import torch
        from gwfast_b200 import waveforms, signal, network, synthetic, _engine, _capi as K
        self.K, self.torch, self.tag = K, torch, tag
        self.ev = ev
        self.n = len(ev['Mc'])
        self.wf = getattr(waveforms, model_name)()
        self.sigs = synthetic.build_network(signal.GWSignal, self.wf, netname, useEarthMotion=rot, fmin=fmin)
        self.net = network.DetNet(self.sigs, verbose=False)
        st = self.st = _engine.state()
        self.lib = st.lib
        self.stream = torch.cuda.current_stream(st.device)
        self.sp = C.c_void_p(self.stream.cuda_stream)
        self.model = self.wf._descriptor(ev)
        dets = [s._detector_struct(i) for i, s in enumerate(self.sigs.values())]
        handles = [s._psd_handle() for s in self.sigs.values()]
        self.ndet, self.npsd = len(dets), len(handles)
        self.darr, self.parr = _engine._call_arrays(dets, handles)
        self.dev_ev, self.host_ev, self.evs, _ = _engine._upload(st, signal._engine_events(self.wf, ev), self.n, K.EVENT_KEYS)
        self.nP = self.lib.gwf_num_params(C.byref(self.model))
        self.npack = self.nP * (self.nP + 1) // 2
        self.narms = self.lib.gwf_num_arms(self.darr, self.ndet)
        n, dev, f64 = self.n, st.device, torch.float64
        self.ws = torch.empty(int(self.lib.gwf_workspace_bytes(C.byref(self.model), n)), dtype=torch.uint8, device=dev)
        self.packed = torch.empty((n, self.npack), dtype=f64, device=dev)
        self.snr2 = torch.empty((n,), dtype=f64, device=dev)
        self.status = torch.empty((n,), dtype=torch.int32, device=dev)
        self.snr2_arm = torch.empty((self.narms, n), dtype=f64, device=dev)
        self.full = torch.empty((self.nP, self.nP, n), dtype=f64, device=dev)
        self.fo = K.gwf_fisher_out(self.packed.data_ptr(), None, self.snr2.data_ptr(), None, self.status.data_ptr())
        self.opts = K.gwf_opts(RES, 0, 0, 0)
        self.opts_reuse = K.gwf_opts(RES, K.GWF_OPT_REUSE_WORKSPACE, 0, 0)

    def _common(self, opts):
        return (C.byref(self.model), self.darr, self.ndet, self.parr, self.npsd, C.byref(self.evs), self.n, C.byref(opts))

    def fisher(self, reuse=False):
        """prologue + fisher_kernel (SNR^2 of SNRInteg, Fisher, status words); reuse: fisher_kernel alone"""
        self.K.check(self.lib.gwf_fisher_ex(*self._common(self.opts_reuse if reuse else self.opts), C.byref(self.fo), C.c_void_p(self.ws.data_ptr()),
                                            self.ws.numel(), self.sp), 'gwf_fisher_ex')

    def set_peer(self, pg):
        """multi-GPU: the Fisher kernel stores every finished packed row into this rank's slot of every rank's gathered buffer"""
        self.fo.peer_fisher = C.cast(pg.slots(0), C.c_void_p)
        self.fo.npeers = pg.world

    def unpack(self, peer=None):
        if peer is not None:
            peer.unpack_and_scatter(self.packed, self.n, self.nP, self.full, self.n, self.stream)
        else:
            self.K.check(self.lib.gwf_unpack_fisher(C.c_void_p(self.packed.data_ptr()), self.n, self.nP, C.c_void_p(self.full.data_ptr()), self.sp), 'gwf_unpack_fisher')

    def snr(self):
        self.K.check(self.lib.gwf_snr(*self._common(self.opts), C.c_void_p(self.snr2_arm.data_ptr()), C.c_void_p(self.ws.data_ptr()), self.ws.numel(), self.sp),
                     'gwf_snr')

    def release(self):
        for k in ('ws', 'packed', 'snr2', 'status', 'snr2_arm', 'full', 'dev_ev'):
            setattr(self, k, None)


def run_engine(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from gwfast_b200 import synthetic, parallel, _engine, _capi as K

    dev = torch.device('cuda', local)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)          # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def timed(fn, steps, sync_ranks=True):
        """sum of the device times of `steps` calls of fn (CUDA events on the launching stream, L2 flushed before each)"""
        stream = torch.cuda.current_stream(dev)
        tot = 0.0
        for _ in range(steps):
            flush.fill_(1)
            if sync_ranks:
                barrier()
            else:
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            if sync_ranks:
                barrier()
            else:
                torch.cuda.synchronize()
            tot += e0.elapsed_time(e1) * 1e-3
        return tot

    # ---- headline: weak scaling, every rank owns its own 10^4-event shard of a world*10^4-event catalog
    full_cat = synthetic.bbh_catalog(EVENTS_PER_GPU * world, synthetic.SEEDS['C2'] if world == 1 else synthetic.SEEDS['C5'])
    n = EVENTS_PER_GPU
    ev = {k: np.ascontiguousarray(v[rank * n:(rank + 1) * n]) for k, v in full_cat.items()}
    case = Case('C2', 'IMRPhenomD', 'ET+2CE', ev)
    net = case.net
    peer = None
    gathered = None
    gather_kind = 'none (one GPU)'
    if world > 1:
        gather_kind = 'dist.all_gather_into_tensor of the packed Fisher (NCCL)'
        if not args.nccl_gather:
            pg = parallel.PeerGather(n, case.npack, dist)
            if pg.available():
                peer = pg
                gathered = pg.gathered
                case.set_peer(pg)
                gather_kind = 'fused into fisher_kernel: every finished packed row is stored into the peers\' gathered buffers over NVLink (gwf_fisher_out.peer_fisher, CUDA IPC peer memory)'
        if peer is None:
            gathered = torch.empty((world, n, case.npack), dtype=torch.float64, device=dev)

    def kernel_step():
        case.fisher()
        case.unpack()
        if world > 1 and peer is None:
            dist.all_gather_into_tensor(gathered.view(-1), case.packed.view(-1))
        return 3                                                                    # prologue, fisher, unpack(+gather)

    def separate_step():
        case.snr()
        case.fisher()
        case.unpack()
        return 5

    sampler = ClockSampler(local)
    sampler.start()

    warm = max(3, args.warmup)
    for _ in range(warm):
        kernel_step()
    barrier()
    t_dev = timed(kernel_step, args.steps)
    launches = 3 * args.steps
    # ---- dominant kernel alone (fisher_kernel): records are still in the workspace
    t_main = timed(lambda: case.fisher(reuse=True), args.steps, sync_ranks=False)
    # ---- the step as two calls (round 1's definition), this rank only
    for _ in range(2):
        separate_step()
    t_sep = timed(separate_step, max(3, args.steps // 4), sync_ranks=False) / max(3, args.steps // 4)

    # ---- multi-GPU result check: 64 events of the next rank's shard, recomputed here, against the gathered rows
    mgc = None
    if world > 1:
        kernel_step()
        barrier()
        other = (rank + 1) % world
        sub = {k: np.ascontiguousarray(v[other * n:other * n + 64]) for k, v in full_cat.items()}
        chk = Case('chk', 'IMRPhenomD', 'ET+2CE', sub)
        chk.fisher()
        torch.cuda.synchronize()
        same = bool(torch.equal(chk.packed, gathered[other, :64])) and bool(torch.equal(gathered[rank], case.packed))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mgc = 'bitwise' if int(flag.item()) == 1 else 'MISMATCH'
        chk.release()

    # ---- end to end through the public API, host numpy in / host numpy out
    def api_step():
        if world > 1:
            # host arrays in, this rank's host arrays out, plus the final gather of the results taken from the engine's
            # device-resident Fisher matrices: the full matrix stays in HBM on every rank
            return parallel.fisher_with_device_gather(net, dict(ev), n * world, dist, peer=peer, res=RES, return_SNR=True)
        return net.FisherMatr(dict(ev), res=RES, return_SNR=True)

    for _ in range(warm):
        # results are held like in the timed loop, so the pinned-buffer pool reaches its steady state here
        r_ = api_step()
    barrier()
    t_e2e = 0.0
    for _ in range(args.steps):
        flush.fill_(1)
        barrier()
        t = time.perf_counter()
        r_ = api_step()
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t
    F_ = r_[0][0] if world > 1 else r_[0]
    h2d = 13 * n * 8
    d2h = (F_.size + n) * 8 + n * 4
    t_e2e_sep = 0.0
    if world == 1:
        for i in range(2 + max(3, args.steps // 4)):
            flush.fill_(1)
            torch.cuda.synchronize()
            t = time.perf_counter()
            s_ = net.SNR(dict(ev), res=RES)
            F2_ = net.FisherMatr(dict(ev), res=RES)
            torch.cuda.synchronize()
            if i >= 2:
                t_e2e_sep += (time.perf_counter() - t) / max(3, args.steps // 4)
    sampler.stop_flag = True
    time.sleep(0.15)

    # ---- measured FP64 peak, same run
    peak = C.c_double(0.)
    K.check(case.lib.gwf_fp64_peak(50.0, C.byref(peak), case.sp), 'gwf_fp64_peak')
    peak_tf = float(peak.value)

    t_dev, t_main, t_e2e = max_over_ranks(t_dev, t_main, t_e2e)
    total_events = n * world * args.steps
    value = total_events / t_dev
    e2e_value = total_events / t_e2e
    achieved = FLOP_PER_EVENT * n * args.steps / t_main / 1e12
    case.release()

    # ---- the other configurations
    others = {}
    if not args.no_other_configs:
        osteps = max(3, min(args.steps, 5))
        specs = [('C1', 'TaylorF2_RestrictedPN', 'ETSL', 'bns', False, True, 2., EVENTS_PER_GPU * world, 'weak',
                  'TaylorF2_RestrictedPN BNS, one L (ETSL), 10^4 events/GPU'),
                 ('C3', 'IMRPhenomD_NRTidalv2', 'ET+2CE', 'bns', True, True, 2., EVENTS_PER_GPU * world, 'weak',
                  'IMRPhenomD_NRTidalv2 BNS, ET+2CE, 13 parameters, 10^4 events/GPU'),
                 ('C4', 'IMRPhenomHM', 'LVK-O4', 'bbh', False, False, 10., 100000, 'strong',
                  'IMRPhenomHM BBH, H1+L1+Virgo+KAGRA (O4), 10^5 events split over the ranks'),
                 ('C5', 'IMRPhenomD', 'ET+2CE', 'bbh', False, True, 2., 1000000, 'strong',
                  '10^6 IMRPhenomD BBH events, ET+2CE, split over the ranks (north-star target; IMRPhenomXAS is not in the reference)'),
                 ('NSBH', 'IMRPhenomNSBH', 'ET+2CE', 'nsbh', True, True, 2., EVENTS_PER_GPU * world, 'weak',
                  'IMRPhenomNSBH, ET+2CE, 13 parameters, 10^4 events/GPU (not a BASELINE.json configuration: no frozen FLOP model, no fraction)')]
        for tag, mname, netname, kind, tidal, rot, fmin, n_tot, scaling, desc in specs:
            cat = (synthetic.bbh_catalog(n_tot, synthetic.SEEDS[tag]) if kind == 'bbh' else
                   synthetic.nsbh_catalog(n_tot, synthetic.SEEDS[tag]) if kind == 'nsbh' else synthetic.bns_catalog(n_tot, synthetic.SEEDS[tag], tidal=tidal))
            lo, hi = parallel.shard_bounds(n_tot, world, rank)
            sub = {k: np.ascontiguousarray(v[lo:hi]) for k, v in cat.items()}
            m = hi - lo
            c = Case(tag, mname, netname, sub, rot=rot, fmin=fmin)
            pgc = None
            if world > 1 and peer is not None:
                nmax = max(parallel.shard_bounds(n_tot, world, r)[1] - parallel.shard_bounds(n_tot, world, r)[0] for r in range(world))
                pgc = parallel.PeerGather(nmax, c.npack, dist)
                if not pgc.available():
                    pgc = None
            gat = torch.empty((world, m, c.npack), dtype=torch.float64, device=dev) if (world > 1 and pgc is None and scaling == 'weak') else None

            if pgc is not None:
                c.set_peer(pgc)

            def kstep():
                c.fisher()
                c.unpack()
                if gat is not None:
                    dist.all_gather_into_tensor(gat.view(-1), c.packed.view(-1))

            for _ in range(2):
                kstep()
            barrier()
            tk = timed(kstep, osteps)
            tm = timed(lambda: c.fisher(reuse=True), osteps, sync_ranks=False)

            def astep():
                if world > 1:
                    return parallel.fisher_with_device_gather(c.net, dict(sub), n_tot, dist, peer=pgc, res=RES, return_SNR=True)
                return c.net.FisherMatr(dict(sub), res=RES, return_SNR=True)

            for _ in range(2):
                r_ = astep()
            barrier()
            te = 0.0
            for _ in range(osteps):
                flush.fill_(1)
                barrier()
                t = time.perf_counter()
                r_ = astep()
                torch.cuda.synchronize()
                te += time.perf_counter() - t
            ok = bool(np.all(np.isfinite(r_[0][0] if world > 1 else r_[0])))
            r_ = None
            tk, tm, te = max_over_ranks(tk, tm, te)
            ach = FLOP_MODEL[tag] * m * osteps / tm / 1e12 if tag in FLOP_MODEL else None
            others[tag] = dict(workload=desc, events_total=n_tot, events_per_gpu=m, scaling=scaling, steps=osteps, value=n_tot * osteps / tk, unit='events/s',
                               ms_per_step=1e3 * tk / osteps, e2e=n_tot * osteps / te, fisher_kernel_ms=1e3 * tm / osteps,
                               fisher_kernel_ms_per_1e4=1e3 * tm / osteps * 1e4 / m, flop_per_event=FLOP_MODEL.get(tag), achieved_tflops=ach,
                               frac=ach / peak_tf if (ach is not None and peak_tf > 0) else None,
                               frac_of_nominal=ach / FP64_NOMINAL_TFLOPS if ach is not None else None, finite=ok,
                               gather=('fused into the Fisher kernel (NVLink peer stores)' if pgc is not None else ('NCCL all-gather' if gat is not None else 'none'))
                               if world > 1 else 'none (one GPU)')
            c.release()
            if pgc is not None:
                pgc.close()
            c = pgc = gat = None
            torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)
        runs = [pool.run(96 * cores, offset=i * 96 * cores) for i in range(3)]   # bounded sample: 3 x ~4 s of CPU work on all host cores
        pool.close()
        rates = sorted(ns / dt for ns, dt in runs)
        cpu = dict(value=rates[1], unit='events/s', cores=cores, kind='port',
                   sample='median of 3 x %d events of the C2 catalog (%.1f s wall in total), multiprocessing.Pool(%d), numpy+dual oracle port' % (
                       runs[0][0], sum(dt for _, dt in runs), cores), runs=rates)
    if rank == 0:
        line = dict(metric='fisher_events_per_s', value=value, unit='events/s', n_gpus=world, steps=args.steps, warmup=warm,
                    ms_per_step=1e3 * t_dev / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                    config=dict(workload=WORKLOAD, events_per_gpu=n, res=RES, l2='flushed between timed iterations (256 MiB fill)',
                                step='SNR + Fisher of every event from one fused launch (gwf_fisher_ex with snr2_integ; API: DetNet.FisherMatr(return_SNR=True))',
                                parallelism='events sharded contiguously, one rank per GPU', gather=gather_kind),
                    e2e=dict(value=e2e_value, unit='events/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches,
                    roofline=dict(bound='fp64', kernel='fisher_kernel<IMRPhenomD,NT=4>', achieved=achieved, peak=peak_tf, unit='TFLOP/s',
                                  frac=achieved / peak_tf if peak_tf > 0 else None, traffic=DRAM_BYTES_PER_LAUNCH,
                                  traffic_source='dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture profiles/r02d_fisher_d_ncu_raw.csv (records + EventAux + PSD windows + events in, packed Fisher out)',
                                  peak_source='measured in this run (gwf_fp64_peak DFMA chain); MEASURED_PEAKS.json has no FP64 entry',
                                  nominal_peak=FP64_NOMINAL_TFLOPS, frac_of_nominal=achieved / FP64_NOMINAL_TFLOPS,
                                  flop_per_event=FLOP_PER_EVENT, kernel_ms=1e3 * t_main / args.steps),
                    separate_calls=dict(kernel_path_ms=1e3 * t_sep, kernel_path_events_per_s=n / t_sep, launches_per_step=5,
                                        e2e_events_per_s=(n / t_e2e_sep) if t_e2e_sep > 0 else None,
                                        note='gwf_snr + gwf_fisher + unpack / DetNet.SNR + DetNet.FisherMatr on this rank (round 1\'s step)'),
                    clocks=sampler.summary())
        if mgc is not None:
            line['multi_gpu_check'] = mgc
        if others:
            line['other_configs'] = others
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        if peer is not None:
            gathered = None
            peer.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--nccl-gather', action='store_true', help='N>1: use dist.all_gather_into_tensor instead of the NVLink peer stores')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
