"""Catalog driver over the engine: the per-batch pipeline of the reference's
``run/calculate_forecasts_from_catalog.py`` -- SNRs, detection threshold, duty factor, Fisher matrices of the detected events,
fixed parameters, condition numbers, covariances, inversion errors, 90 % sky areas (``compute_errs``, run script :410-634) --
and its chunking / resume logic (:738, :900-1007), with every per-event computation on the GPU
(``DetNet.SNR`` / ``DetNet.FisherMatr`` -> fisher/snr kernels, ``fisherTools`` -> covariance kernels).

Differences from the reference script, on purpose: batches are sized for a GPU (1e5 events by default, not 1 per CPU pool
worker), ranks of a ``torch.distributed`` job take batches round-robin instead of a multiprocessing pool, and batch files are
``.npz`` (HDF5 only if ``h5py`` is installed).  Plotting, logging tees and the LAL waveform options are not provided.
"""
import argparse
import copy
import json
import os
import time

import numpy as np

from . import fisherTools as ft

SAVE_KEYS = ('snrs', 'errors', 'sky_area_90', 'cond_numbers', 'eps', 'idxs_detected')


def get_events_subset(events, mask):
    """run script :398-403."""
    return {k: np.asarray(v)[mask] for k, v in events.items()}


def compute_errs(events, net, snr_th=12., duty_factor=None, seeds=None, params_fix=(), compute_fisher=True, return_all=True,
                 return_derivatives=False, return_snr_derivatives=False, i_in=0, i_f=None, fisher_kwargs=None):
    """One batch: returns ``(snrs_all, Fres, [derivatives_all,] eps_dL, Cov_dL, sky_area_90, cond_numbers, idxs_detected)`` exactly
    like the reference's compute_errs (run script :410-634)."""
    nevents = len(events[list(events.keys())[0]])
    i_f = i_in + nevents if i_f is None else i_f
    fkw = dict(res=1000, df=None, spacing='geom', use_chi1chi2=True, computeAnalyticalDeriv=True)
    fkw.update(fisher_kwargs or {})
    snrs_df1 = net.SNR(events, return_all=True)
    no_duty = duty_factor is None or duty_factor >= 1
    snrs_all = copy.deepcopy(snrs_df1)
    applied = {}
    if not no_duty:
        # one Bernoulli mask per arm, seeded per detector as in the run script (:423-448)
        net2 = np.zeros(nevents)
        for i, key in enumerate(net.signals):
            np.random.seed(None if seeds is None else seeds[i])
            arms = [key] if net.signals[key].detector_shape == 'L' else [key + '_%s' % a for a in range(3)]
            for a in arms:
                applied[a] = np.random.choice([0, 1], nevents, p=[1. - duty_factor, duty_factor])
                snrs_all[a] = snrs_all[a] * applied[a]
                net2 += snrs_all[a] ** 2
        snrs_all['net'] = np.sqrt(net2)
    snrs = snrs_all['net']
    detected = np.atleast_1d(snrs > snr_th)
    idxs_detected = np.arange(i_in, i_f)[np.argwhere(detected)]
    events_det = get_events_subset(events, detected)
    ndet = int(detected.sum())
    wf = net.signals[list(net.signals.keys())[0]].wf_model
    npar = wf.nParams - len(params_fix)
    shape = (npar, npar, ndet)
    want_d = return_derivatives or return_snr_derivatives
    derivatives_all = {}
    if not compute_fisher or ndet == 0:
        nanF = np.full(shape, np.nan)
        out = (snrs_all, nanF) + ((derivatives_all,) if want_d else ()) + \
              (np.full(nevents, np.nan), np.full(shape, np.nan), np.full(nevents, np.nan), np.full(nevents, np.nan), idxs_detected)
        return out
    if want_d:
        Fres_, derivatives_all = net.FisherMatr(events_det, return_all=True, return_derivatives=return_derivatives,
                                                return_SNR_derivatives=return_snr_derivatives, **fkw)
    else:
        Fres_ = net.FisherMatr(events_det, return_all=True, **fkw)
    if not no_duty:
        netF = 0.
        netD = 0.
        for a, m in applied.items():
            md = m[detected]
            Fres_[a] = Fres_[a] * md
            netF = netF + Fres_[a]
            if return_derivatives:
                derivatives_all[a] = derivatives_all[a] * md[None, :, None]
            elif return_snr_derivatives:
                derivatives_all[a] = derivatives_all[a] * md[None, :]
                netD = netD + derivatives_all[a]
        Fres_['net'] = netF
        if return_snr_derivatives and not return_derivatives:
            derivatives_all['net'] = netD / snrs[detected]
    parNums = wf.ParNums
    if len(params_fix) > 0:
        if return_all:
            Fres = {k: ft.fixParams(v, wf.ParNums, list(params_fix))[0] for k, v in Fres_.items()}
            totF, parNums = ft.fixParams(Fres_['net'], wf.ParNums, list(params_fix))
        else:
            totF, parNums = ft.fixParams(Fres_['net'], wf.ParNums, list(params_fix))
            Fres = totF
    else:
        totF, Fres = Fres_['net'], (Fres_ if return_all else Fres_['net'])
    _, _, cond_numbers = ft.CheckFisher(totF)
    Cov_dL, eps_dL = ft.CovMatr(totF, invMethodIn='cho', condNumbMax=1e50, svals_thresh=1e-15, truncate=False, verbose=False)
    sky = ft.compute_localization_region(Cov_dL, parNums, events_det['theta'], perc_level=90, units='SqDeg')
    return (snrs_all, Fres) + ((derivatives_all,) if want_d else ()) + (eps_dL, Cov_dL, sky, cond_numbers, idxs_detected)


# ---------------------------------------------------------------------------------------------- catalog IO
def load_catalog(path):
    """events dict from ``.npz`` (one array per key), ``.txt``/``.dat`` (whitespace table with a ``# key key ...`` header), or the
    reference's HDF5 population file (gwfastUtils.py:70-146) when h5py is installed."""
    ext = os.path.splitext(path)[1].lower()
    if ext == '.npz':
        z = np.load(path)
        return {k: np.asarray(z[k], dtype=float) for k in z.files}
    if ext in ('.txt', '.dat'):
        with open(path) as fh:
            head = fh.readline()
        if not head.startswith('#'):
            raise ValueError('text catalogs need a "# Mc eta dL ..." header line')
        keys = head[1:].split()
        tab = np.atleast_2d(np.loadtxt(path))
        return {k: tab[:, i].copy() for i, k in enumerate(keys)}
    if ext in ('.h5', '.hdf5'):
        try:
            import h5py
        except ImportError:
            raise ImportError('reading %s needs h5py, which is not installed; convert the catalog to .npz' % path)
        with h5py.File(path, 'r') as f:
            return {k: np.array(f[k]) for k in f.keys()}
    raise ValueError('unknown catalog format %s' % ext)


def batch_ranges(n, batch_size, idx_in=0, idx_f=None):
    """contiguous [start, stop) batches of the slice [idx_in, idx_f) (run script :900-1007, one list instead of per-pool lists)."""
    idx_f = n if idx_f is None else min(idx_f, n)
    return [(a, min(a + batch_size, idx_f)) for a in range(idx_in, idx_f, batch_size)]


def _batch_file(fout, a, b):
    return os.path.join(fout, 'batch_%d_to_%d.npz' % (a, b))


REFERENCE_FILES = ('snrs', 'fishers', 'covs', 'sky_area', 'errors', 'inversion_errors', 'cond_numbers', 'idxs_det')


def to_reference_files(fout, a, b, snrs, totF, eps, cov, sky, errs, idxs, cond):
    """One batch in the reference's own files (run script to_file :224-247 and the snrs-only branch :784-788): ``snrs_<a>_to_<b>.txt``
    for every batch -- the file ``--resume_run`` looks for (:738-741) -- and, when Fisher matrices were computed, ``fishers``/``covs``
    ``.npy`` plus ``sky_area``, ``errors``, ``inversion_errors``, ``cond_numbers``, ``idxs_det`` ``.txt`` with the same suffix, so that the
    reference script can resume or concatenate a run started here (and vice versa)."""
    suff = '_%d_to_%d' % (a, b)
    np.savetxt(os.path.join(fout, 'snrs' + suff + '.txt'), np.atleast_1d(snrs))
    if totF is None or np.all(np.isnan(totF)):
        return
    np.save(os.path.join(fout, 'fishers' + suff + '.npy'), totF)
    np.save(os.path.join(fout, 'covs' + suff + '.npy'), cov)
    np.savetxt(os.path.join(fout, 'sky_area' + suff + '.txt'), np.atleast_1d(sky))
    np.savetxt(os.path.join(fout, 'errors' + suff + '.txt'), errs)
    np.savetxt(os.path.join(fout, 'inversion_errors' + suff + '.txt'), np.atleast_1d(eps))
    np.savetxt(os.path.join(fout, 'cond_numbers' + suff + '.txt'), np.atleast_1d(cond))
    np.savetxt(os.path.join(fout, 'idxs_det' + suff + '.txt'), idxs)


def concatenate_reference_files(fout, ranges):
    """The reference's ``--concatenate`` step (run script :1034-1170) over batches written by to_reference_files: one array per quantity
    in catalog order (batches without detections contribute their SNRs only)."""
    out = {k: [] for k in ('snrs', 'fishers', 'covs', 'sky_area', 'errors', 'inversion_errors', 'cond_numbers', 'idxs_det')}
    for a, b in ranges:
        suff = '_%d_to_%d' % (a, b)
        out['snrs'].append(np.atleast_1d(np.loadtxt(os.path.join(fout, 'snrs' + suff + '.txt'))))
        if not os.path.exists(os.path.join(fout, 'fishers' + suff + '.npy')):
            continue
        out['fishers'].append(np.load(os.path.join(fout, 'fishers' + suff + '.npy')))
        out['covs'].append(np.load(os.path.join(fout, 'covs' + suff + '.npy')))
        for k in ('sky_area', 'inversion_errors', 'cond_numbers', 'idxs_det'):
            out[k].append(np.atleast_1d(np.loadtxt(os.path.join(fout, k + suff + '.txt'))))
        e = np.loadtxt(os.path.join(fout, 'errors' + suff + '.txt'))
        out['errors'].append(e[:, None] if e.ndim == 1 else e)
    res = {'snrs': np.concatenate(out['snrs'])}
    if out['fishers']:
        res.update(fishers=np.concatenate(out['fishers'], axis=-1), covs=np.concatenate(out['covs'], axis=-1), errors=np.concatenate(out['errors'], axis=-1))
        for k in ('sky_area', 'inversion_errors', 'cond_numbers', 'idxs_det'):
            res[k] = np.concatenate(out[k])
    return res


def run_catalog(events, net, fout, batch_size=100000, snr_th=12., duty_factor=None, seeds=None, params_fix=(), resume=False,
                idx_in=0, idx_f=None, rank=0, world=1, save_fishers=False, verbose=True, reference_files=False):
    """Process a catalog in batches; rank ``rank`` of ``world`` takes every ``world``-th batch.  Each batch is written to
    ``fout/batch_<a>_to_<b>.npz`` as soon as it is done, ``resume`` skips batches whose file exists (run script :738).  With
    ``reference_files`` the batches are written in the reference's own layout instead (``snrs_<a>_to_<b>.txt``, ``fishers_...npy``, ...:
    to_reference_files) and ``resume`` looks for ``snrs_<a>_to_<b>.txt`` exactly like the reference's ``--resume_run``.  Returns the list
    of batch files of this rank."""
    os.makedirs(fout, exist_ok=True)
    n = len(events[list(events.keys())[0]])
    wf = net.signals[list(net.signals.keys())[0]].wf_model
    done = []
    for bi, (a, b) in enumerate(batch_ranges(n, batch_size, idx_in, idx_f)):
        if bi % world != rank:
            continue
        path = os.path.join(fout, 'snrs_%d_to_%d.txt' % (a, b)) if reference_files else _batch_file(fout, a, b)
        if resume and os.path.exists(path):
            done.append(path)
            continue
        t0 = time.time()
        sub = {k: np.asarray(v)[a:b] for k, v in events.items()}
        snrs_all, Fres, eps, cov, sky, cond, idxs = compute_errs(sub, net, snr_th=snr_th, duty_factor=duty_factor, seeds=seeds,
                                                                 params_fix=params_fix, i_in=a, i_f=b)
        errs = np.sqrt(np.einsum('iin->in', cov)) if cov.size else np.zeros((cov.shape[0], 0))
        if reference_files:
            totF = Fres['net'] if isinstance(Fres, dict) else Fres
            to_reference_files(fout, a, b, snrs_all['net'], totF, eps, cov, sky, errs, idxs, cond)
            done.append(path)
            if verbose:
                print('[rank %d] events %d-%d: %d detected (SNR > %s) in %.2f s' % (rank, a, b, len(np.ravel(idxs)), snr_th, time.time() - t0))
            continue
        data = dict(snrs=snrs_all['net'], errors=errs, sky_area_90=sky, cond_numbers=cond, eps=eps, idxs_detected=np.ravel(idxs),
                    par_names=np.array([k for k in wf.ParNums if k not in params_fix]))
        for k, v in snrs_all.items():
            data['snr__' + k] = v
        if save_fishers:
            data['fisher_net'] = Fres['net'] if isinstance(Fres, dict) else Fres
            data['cov'] = cov
        tmp = path + '.tmp.npz'
        np.savez(tmp, **data)
        os.replace(tmp, path)
        done.append(path)
        if verbose:
            print('[rank %d] events %d-%d: %d detected (SNR > %s) in %.2f s' % (rank, a, b, len(np.ravel(idxs)), snr_th, time.time() - t0))
    return done


def collect(fout):
    """Concatenate the batch files of a finished run (all ranks) in catalog order."""
    files = sorted((f for f in os.listdir(fout) if f.startswith('batch_') and f.endswith('.npz') and '.tmp' not in f),
                   key=lambda s: int(s.split('_')[1]))
    parts = [np.load(os.path.join(fout, f)) for f in files]
    if not parts:
        raise FileNotFoundError('no batch files in %s' % fout)
    out = {'snrs': np.concatenate([p['snrs'] for p in parts]), 'idxs_detected': np.concatenate([p['idxs_detected'] for p in parts])}
    for k in ('sky_area_90', 'cond_numbers', 'eps'):
        out[k] = np.concatenate([np.atleast_1d(p[k])[:len(p['idxs_detected'])] if len(p['idxs_detected']) else np.zeros(0) for p in parts])
    out['errors'] = np.concatenate([p['errors'] if len(p['idxs_detected']) else np.zeros((p['errors'].shape[0], 0)) for p in parts], axis=1)
    out['par_names'] = parts[0]['par_names']
    return out


def main(argv=None):
    from . import waveforms, signal, network, synthetic
    ap = argparse.ArgumentParser(description='SNRs, Fisher errors and sky areas for a catalog on the B200 engine')
    ap.add_argument('--catalog', required=True, help='.npz / .txt / .h5 events file')
    ap.add_argument('--fout', required=True)
    ap.add_argument('--wf_model', default='IMRPhenomD', choices=['NewtInspiral', 'TaylorF2_RestrictedPN', 'tf2', 'tf2_tidal', 'IMRPhenomD', 'IMRPhenomD_NRTidalv2', 'IMRPhenomHM'])
    ap.add_argument('--net', default='ET+2CE', choices=sorted(synthetic.NETWORKS))
    ap.add_argument('--fmin', type=float, default=2.)
    ap.add_argument('--no_rot', action='store_true')
    ap.add_argument('--snr_th', type=float, default=12.)
    ap.add_argument('--duty_factor', type=float, default=None)
    ap.add_argument('--params_fix', nargs='*', default=[])
    ap.add_argument('--batch_size', type=int, default=100000)
    ap.add_argument('--idx_in', type=int, default=0)
    ap.add_argument('--idx_f', type=int, default=None)
    ap.add_argument('--resume_run', action='store_true')
    ap.add_argument('--save_fishers', action='store_true')
    ap.add_argument('--reference_files', action='store_true', help="write the batches in the reference script's own files (snrs_<a>_to_<b>.txt, fishers_...npy, ...)")
    args = ap.parse_args(argv)
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    presets = {'tf2': ('TaylorF2_RestrictedPN', dict(use_3p5PN_SpinHO=True)), 'tf2_tidal': ('TaylorF2_RestrictedPN', dict(use_3p5PN_SpinHO=True, is_tidal=True)),
               'NewtInspiral': ('NewtInspiral', dict(is_chi1chi2=False))}          # run script :64-68
    cls, kw = presets.get(args.wf_model, (args.wf_model, {}))
    wf = getattr(waveforms, cls)(**kw)
    net = network.DetNet(synthetic.build_network(signal.GWSignal, wf, args.net, useEarthMotion=not args.no_rot, fmin=args.fmin), verbose=False)
    events = load_catalog(args.catalog)
    os.makedirs(args.fout, exist_ok=True)
    if rank == 0:
        with open(os.path.join(args.fout, 'config.json'), 'w') as fh:
            json.dump(vars(args), fh, indent=1)
    run_catalog(events, net, args.fout, batch_size=args.batch_size, snr_th=args.snr_th, duty_factor=args.duty_factor, params_fix=tuple(args.params_fix),
                resume=args.resume_run, idx_in=args.idx_in, idx_f=args.idx_f, rank=rank, world=world, save_fishers=args.save_fishers, reference_files=args.reference_files)


if __name__ == '__main__':
    main()
