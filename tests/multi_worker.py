"""Worker of the multi-GPU parity test: run under torchrun, one process per GPU.  Every rank runs the same adjustment
through the distributed C-ABI path; rank 0 compares with the CPU oracle and prints one JSON line.

  torchrun ... tests/multi_worker.py <scene> [dense|structured|auto]
  scenes: example | cfg2 | cfg3 | cfg4 | cfg4mid (100 images x 2 700 targets, n = 8 720: compared with the blocked fast oracle)
  JAICOV_PANEL_TILES=1 makes the block-column panels of the distributed Cholesky 128 columns wide, so that even the small
  scenes run many panels per rank (look-ahead, double-buffered staging, trapezoid updates: cfg2 -> 15 panels, cfg4mid -> 69)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bundle_adjustment_b200 as ba  # noqa: E402
from bundle_adjustment_b200 import verify  # noqa: E402
from oracle.fast_oracle import FastOracle  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from tests.helpers import flat_problem  # noqa: E402
from tests.scenes import example_scene, synthetic_scene  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    which = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    if which == 'example':
        scene = example_scene()
    elif which == 'cfg3':
        scene = synthetic_scene(3, images=12, targets=150)[0]
    elif which == 'cfg4':
        scene = synthetic_scene(4, images=30, targets=300)[0]
    elif which == 'cfg4mid':
        scene = synthetic_scene(4, images=100, targets=2700)[0]
    else:
        scene = synthetic_scene(2)[0]
    adj, flat = flat_problem(scene)
    ids = [ba._lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    solver = {'dense': ba._lib.SOLVER_DENSE, 'structured': ba._lib.SOLVER_STRUCTURED, 'auto': ba._lib.SOLVER_AUTO}[
        sys.argv[2] if len(sys.argv) > 2 else 'auto']
    s = ba.Session(sigma2apriori=adj.getVarianceFactorApriori(), device=local, solver=solver)
    s.dist_init(rank, world, ids[0])
    s.set_problem(flat)
    rc = s.estimate()
    st = s.stats()
    n = s.n
    blk = torch.from_numpy(s.qxx_block(0, n, 0, n)).cuda()
    dist.all_reduce(blk)            # partial blocks: the sum over ranks is Qxx
    # the rank's own column tiles as stored (lower part), against the assembled matrix
    cols_l, blocks = s.qxx_local()
    Qfull = blk.cpu().numpy()
    u = int(flat['n_unknowns'])
    local_err = 0.0
    for c0, b in zip(cols_l, blocks):
        c0 = int(c0)
        w = min(128, n - c0)
        rows = min(b.shape[0], n - c0)
        ref = Qfull[c0:c0 + rows, c0:c0 + w]
        got = b[:rows, :w]
        mask = np.tril(np.ones((rows, w), bool))        # on and below the tile's diagonal
        local_err = max(local_err, float(np.abs(np.where(mask, got - ref, 0.0)).max()))
    le = torch.tensor([local_err], dtype=torch.float64, device='cuda')
    dist.all_reduce(le, op=dist.ReduceOp.MAX)

    def reduce_sum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()
    mem = torch.tensor(s.device_bytes(), dtype=torch.float64, device='cuda')
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)          # the largest rank's buffers
    chk = verify.check_pass(s, columns=verify.sample_columns(n, n - u, world=world, panel=128 * int(os.environ.get('JAICOV_PANEL_TILES', '16'))),
                            reduce_sum=reduce_sum, omega=st.omega, values_updated=True)
    if rank == 0:
        o = (FastOracle if which == 'cfg4mid' else Oracle)(scene)
        so = o.estimate()
        Qo = o.qxx_dense()
        Qg = blk.cpu().numpy()
        d = o.fp.d
        sg = np.sqrt(np.abs(np.diag(Qo)))
        sg[:d] = 1.0
        errq = float((np.abs(Qg - Qo) / np.outer(sg, sg)).max())
        s2o = o.variance_factor_aposteriori()
        xyz, io, coef, eo = s.values()
        errx = 0.0
        for vg, vo, cols in ((xyz, o.fp.xyz, o.fp.pt_col), (io, o.fp.io_val, o.fp.io_col), (coef, o.fp.coef_val, o.fp.coef_col),
                             (eo, o.fp.eo_val, o.fp.eo_col)):
            c = cols.astype(np.int64)
            act = (c >= 0) & (c < 2147483647)
            floor = np.sqrt(s2o * np.abs(np.diag(Qo))[c[act]])
            errx = max(errx, float((np.abs(vg[act] - vo[act]) / np.maximum(np.abs(vo[act]), floor)).max()))
        print(json.dumps({'scene': which, 'world': world, 'solver_used': st.solver_used, 'rc': rc, 'rc_oracle': so, 'iterations': st.iterations,
                          'iterations_oracle': len(o.history), 'sigma2_rel_err': abs(st.sigma2aposteriori - s2o) / s2o,
                          'qxx_scaled_err': errq, 'param_rel_err': errx, 'qxx_local_vs_block_maxabs': float(le[0]), 'ms_last_pass': st.ms_total,
                          'ms_factor': st.ms_factor, 'ms_inverse': st.ms_inverse, 'n': n, 'd': n - u, 'panel_tiles': int(os.environ.get('JAICOV_PANEL_TILES', '16')),
                          'device_bytes_max': [int(v) for v in mem.cpu().tolist()], 'verify': {k: chk[k] for k in ('datum_residual', 'cofactor_residual', 'omega_rel_diff')}}), flush=True)
    s.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
