#!/usr/bin/env python
"""Smallest possible first contact of the int8-digit tcgen05 kernel with a B200 (a few seconds of GPU time): three tile-grid
products through jaicov_gemm_tiles with JAICOV_GEMM_OZAKI=8, compared with numpy, every line flushed to gpurun_out/ozaki_quick.log
so that a kill (hang) still leaves what was learnt.  tools/ozaki_gpu_check.py is the full check."""
import os
import sys
import time

t0 = time.time()
os.environ['JAICOV_GEMM_OZAKI'] = os.environ.get('JAICOV_GEMM_OZAKI', '8')
os.environ['JAICOV_OZAKI_MIN_TILES'] = '1'
os.environ['JAICOV_OZAKI_MIN_K'] = '128'
import numpy as np   # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
log = open(os.path.join(ROOT, 'gpurun_out', 'ozaki_quick.log'), 'a')


def P(*a):
    s = ' '.join(str(x) for x in a)
    print(s, flush=True)
    log.write(s + '\n')
    log.flush()
    os.fsync(log.fileno())


P('--- ozaki_quick, digits', os.environ['JAICOV_GEMM_OZAKI'], 'cluster', os.environ.get('JAICOV_OZAKI_CLUSTER', '1'), 'kscale', os.environ.get('JAICOV_OZAKI_KSCALE', '1'))
import bundle_adjustment_b200 as ba   # noqa: E402
L = ba._lib.load()
P('library loaded after %.2f s, devices %d' % (time.time() - t0, L.jaicov_device_count()))
rng = np.random.default_rng(0)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ozaki_gpu_check as chk   # noqa: E402


def case(tag, al, bl, mt, nt, K, alpha, beta, tri, kmode):
    A, B, As, Bs = chk.masked_operands(rng, al, bl, mt, nt, K, kmode)      # NaN wherever the kernels must not read
    C0 = rng.standard_normal((128 * mt, 128 * nt))
    n0 = L.jaicov_launch_count()
    P(tag, 'launching ...')
    C, ms = ba._lib.gemm_tiles(As, Bs, C0, al, bl, alpha, beta, tri, kmode)
    launches = L.jaicov_launch_count() - n0
    ref = alpha * (A @ B.T) + beta * C0
    scale = np.abs(A) @ np.abs(B.T) + np.abs(C0)
    E = np.abs(C - ref) / scale
    if tri:                                                                  # tiles above the diagonal are not computed
        it, jt = np.arange(128 * mt)[:, None] // 128, np.arange(128 * nt)[None, :] // 128
        E = np.where(it >= jt, E, 0.0)
    P(tag, 'max err %.3e, entries within 1e-13: %.4f, launches %d (1 = fell back to the FP64 kernel), %.3f ms, t = %.2f s'
      % (np.nanmax(E), float((E < 1e-13).mean()), launches, ms, time.time() - t0))
    if not (np.nanmax(E) < 1e-13):
        bad = ~(E < 1e-13)
        P('  C[0,:4]  ', C[0, :4])
        P('  ref[0,:4]', ref[0, :4])
        P('  C[1,:4]  ', C[1, :4], ' C[8,:4]', C[8, :4])
        P('  ref[1,:4]', ref[1, :4], ' ref[8,:4]', ref[8, :4])
        P('  bad rows (first 24):', np.unique(np.nonzero(bad)[0])[:24].tolist())
        P('  bad cols (first 24):', np.unique(np.nonzero(bad)[1])[:24].tolist())
        P('  finite: %s, zeros: %.3f, |C| max %.3e vs |ref| max %.3e' % (np.isfinite(C).all(), float((C == 0).mean()), np.nanmax(np.abs(C)), np.abs(ref).max()))
        # is C the right product with rows / columns permuted?  best-matching reference row for C rows 0, 1, 8, 64
        for r in (0, 1, 8, 64):
            d = np.abs(ref - C[r][None, :]).sum(axis=1)
            P('  C row %d is closest to ref row %d (distance %.3e)' % (r, int(np.argmin(d)), d.min()))
        dcol = np.abs(ref[0][None, :].T - C[0][None, :]).argmin(axis=0)[:16]
        P('  C[0, j] closest to ref[0, .] at columns', dcol.tolist())


case('A  1x1 tile, K=128, al=0 bl=0', 0, 0, 1, 1, 128, 1.0, 0.0, 0, 0)
case('B  2x2 tiles, K=512, al=0 bl=1, beta=1', 0, 1, 2, 2, 512, -1.0, 1.0, 0, 0)
case('C  3x3 lower, K=384, al=1 bl=1, k>=max(i,j)', 1, 1, 3, 3, 384, 1.0, 0.0, 1, 3)
P('done after %.2f s' % (time.time() - t0))
