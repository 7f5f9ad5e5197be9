"""GPU tests of the multi-rank path on ONE device: P contexts (one per part) exchange their halo blocks through
the same buffers NCCL would fill; the assembled RHS must be bit-identical to the single-context run
(redundant cut faces + canonical orientation; SURVEY 8e 'determinism'), the VJP equal to ~1e-15."""
import numpy as np
import pytest

import _pkg
from tests import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hg():
    return _pkg.load()


def _route(ctxs, infos, with_lambda):
    """Copy every rank's send block k into the matching recv block of its peer (what send/recv does)."""
    import torch
    from hydrograd_jl_b200.parallel import _DevArray
    bufs = []
    for c in ctxs:
        c.halo_pack(with_lambda)
        c.sync()
        sp, rp, n = c.halo_buffers()
        bufs.append((torch.as_tensor(_DevArray(sp, n), device="cuda") if n else None,
                     torch.as_tensor(_DevArray(rp, n), device="cuda") if n else None))
    per = 6 if with_lambda else 3
    for r, info in enumerate(infos):
        off = np.concatenate([[0], np.cumsum([6 * c for c in info["counts"]])])
        for k, (peer, nk) in enumerate(zip(info["neighbors"], info["counts"])):
            pinfo = infos[peer]
            kk = pinfo["neighbors"].index(r)
            poff = int(np.concatenate([[0], np.cumsum([6 * c for c in pinfo["counts"]])])[kk])
            bufs[peer][1][poff:poff + per * nk] = bufs[r][0][int(off[k]):int(off[k]) + per * nk]
    torch.cuda.synchronize()


@pytest.mark.parametrize("case", ["dam_rcb4", "river_slab3", "dam_thin_rcb3", "river_rcb4_inlet"])
def test_partitioned_rhs_and_vjp_match_single_context_bitwise(hg, case):
    from hydrograd_jl_b200 import parallel as P
    from hydrograd_jl_b200 import synthetic as S
    if case == "dam_rcb4":
        flat, Q0 = S.dam_break(40); Pn = 4
    elif case == "dam_thin_rcb3":
        flat, Q0 = S.dam_break(36, thin_film=True); Pn = 3
    elif case == "river_rcb4_inlet":
        flat, Q0 = S.river(60, 48); Pn = 4      # RCB cuts through the inlet node-string: its cells are kept on one rank
    else:
        flat, Q0 = S.river(90, 24); Pn = 3
    N = flat["n_cells"]
    if case == "river_slab3":
        part = (np.arange(N) * Pn // N).astype(np.int32)
    elif case == "river_rcb4_inlet":
        cx, cy = flat["cell_centroids"][:N], flat["cell_centroids"][N:]
        groups = P.inlet_cell_groups(flat)
        assert len(np.unique(P.rcb_partition(cx, cy, Pn)[groups[0]])) > 1, "plain RCB must split the inlet in this case"
        part = P.rcb_partition(cx, cy, Pn, keep_together=groups)
    else:
        part = P.rcb_partition(flat["cell_centroids"][:N], flat["cell_centroids"][N:], Pn)
    rng = np.random.default_rng(2)
    Q = cases.random_state_flat(flat, 7, dry_frac=0.05) if not case.startswith("river") else Q0
    lam = rng.standard_normal(3 * N)
    single = hg.Context(flat, tile_cells=128)
    ref = single.rhs(Q)
    ref_bar, _ = single.rhs_vjp(Q, lam)
    locs = [P.extract_local(flat, part, r, Q) for r in range(Pn)]
    ctxs = [hg.Context(loc, tile_cells=128) for loc, _ in locs]
    infos = [info for _, info in locs]
    for c, info in zip(ctxs, infos):
        c.set_state(info["Q"])
        c.set_lambda(np.concatenate([lam[k * N + info["own"]] for k in range(3)]))
    _route(ctxs, infos, with_lambda=True)
    got = np.zeros(3 * N)
    bar = np.zeros(3 * N)
    for c, info in zip(ctxs, infos):
        # overlap launches first (fresh output buffers): tiles without halo faces, then the band; same bits as one launch
        c.rhs_resident(1); c.rhs_resident(2)
        d12 = c.get_rhs()
        c.vjp_resident(1); c.vjp_resident(2)
        b12, _ = c.get_vjp()
        c.rhs_resident()
        d = c.get_rhs()
        c.vjp_resident()
        b, _ = c.get_vjp()
        assert np.array_equal(d12, d) and np.array_equal(b12, b)
        n = info["own"].size
        for k in range(3):
            got[k * N + info["own"]] = d[k * n:(k + 1) * n]
            bar[k * N + info["own"]] = b[k * n:(k + 1) * n]
    assert np.array_equal(got, ref), f"max diff {np.abs(got - ref).max()}"
    # the adjoint of a cut face runs through the boundary-face copy of the sweep: same arithmetic, last-place differences
    assert np.abs(bar - ref_bar).max() <= 1e-13 * np.abs(ref_bar).max(), f"max diff {np.abs(bar - ref_bar).max()}"


@pytest.mark.parametrize("case", ["dam_rcb4", "river_slab3"])
def test_library_owned_exchange_matches_single_context(hg, case):
    """The library's own halo transport (hg_comm.cu: peer stores + epoch flags, band tiles of the ONE launch wait) between
    rank contexts of this process: RHS bit-identical to the single context, VJP to rounding, Euler steps (every step an
    exchange, two parity buffers) bit-identical after 6 steps.  On one device all ranks must push before any consumes
    (auto mode off); the multi-process, multi-GPU run of the same path is tests/test_gpu_torchrun.py."""
    from hydrograd_jl_b200 import parallel as P
    from hydrograd_jl_b200 import synthetic as S
    if case == "dam_rcb4":
        flat, Q0 = S.dam_break(40); Pn = 4
        N = flat["n_cells"]
        part = P.rcb_partition(flat["cell_centroids"][:N], flat["cell_centroids"][N:], Pn)
        Q = cases.random_state_flat(flat, 7, dry_frac=0.05)
    else:
        flat, Q0 = S.river(90, 24); Pn = 3
        N = flat["n_cells"]
        part = (np.arange(N) * Pn // N).astype(np.int32)
        Q = Q0
    lam = np.random.default_rng(3).standard_normal(3 * N)
    single = hg.Context(flat, tile_cells=128)
    ref = single.rhs(Q)
    ref_bar, _ = single.rhs_vjp(Q, lam)
    locs = [P.extract_local(flat, part, r, Q) for r in range(Pn)]
    ctxs = [hg.Context(loc, tile_cells=128) for loc, _ in locs]
    infos = [info for _, info in locs]
    P.connect_contexts(ctxs, infos)
    for c, info in zip(ctxs, infos):
        c.set_state(info["Q"])
        c.set_lambda(np.concatenate([lam[k * N + info["own"]] for k in range(3)]))
        c.comm_set_auto(False)

    def assemble(getter):
        out = np.zeros(3 * N)
        for c, info in zip(ctxs, infos):
            d = getter(c)
            n = info["own"].size
            for k in range(3):
                out[k * N + info["own"]] = d[k * n:(k + 1) * n]
        return out

    # without an exchange the evaluation refuses (auto mode off)
    with pytest.raises(hg.HydrogradError):
        ctxs[0].rhs_resident()
    for c in ctxs:
        c.comm_exchange(False)
    for c in ctxs:
        c.rhs_resident()
    assert np.array_equal(assemble(lambda c: c.get_rhs()), ref)
    for c in ctxs:
        c.comm_exchange(True)
    for c in ctxs:
        c.vjp_resident()
    bar = assemble(lambda c: c.get_vjp()[0])
    assert np.abs(bar - ref_bar).max() <= 1e-13 * np.abs(ref_bar).max()
    # time stepping: one exchange per step, alternating parity buffers
    dt = 1e-3
    single.set_state(Q)
    single.step_euler(dt, 6)
    for _ in range(6):
        for c in ctxs:
            c.comm_exchange(False)
        for c in ctxs:
            c.step_euler(dt, 1)
    assert np.array_equal(assemble(lambda c: c.get_state()), single.get_state())
    # a consumer whose neighbours never pushed gives up with HG_ERR_COMM instead of hanging the device
    ctxs[0].comm_exchange(False)
    with pytest.raises(hg.HydrogradError, match="COMM|halo exchange"):
        ctxs[0].rhs_resident()
        ctxs[0].get_rhs()
