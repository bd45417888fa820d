"""One rank of a multi-process slab-decomposed run (launched by tests/test_multirank_gloo.py on the
CPU emulator, and by tests/test_gpu_multirank.py on real GPUs): wires the ranks exactly as
bench.py does (all-gather of the peer blobs through torch.distributed), advances the same initial
data as the single-grid oracle and compares this rank's slabs with the oracle's."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parity_common as pc  # noqa: E402
from laps_b200 import Solver  # noqa: E402
from oracle import laps_oracle as lo  # noqa: E402


def connect(g, world, device=None):
    blob = torch.from_numpy(np.frombuffer(g.export_peer_blob(), dtype=np.uint8).copy())
    if device is not None:
        blob = blob.to(device)
    blobs = [torch.empty_like(blob) for _ in range(world)]
    dist.all_gather(blobs, blob)
    g.import_peer_blobs(b"".join(b.cpu().numpy().tobytes() for b in blobs))
    dist.barrier()


def main():
    cfg = json.loads(sys.argv[1])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend = cfg.get("backend", "gloo")
    device = None
    if backend == "nccl":
        torch.cuda.set_device(rank)
        device = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=device)
    else:
        dist.init_process_group("gloo")
    lib = cfg.get("lib")
    nx, ny, nz = cfg["shape"]
    make = pc.make_case_incompressible if cfg.get("incompressible") else pc.make_case
    p, prim = make(nx, ny, nz, **cfg.get("case", {}))
    g = Solver(lib, rank=rank, nranks=world, device=rank if backend == "nccl" else 0, **pc.solver_kwargs(p))
    connect(g, world, device)
    zo, zn, yo, yn = g.ext.z_offset, g.ext.z_size, g.ext.y_offset, g.ext.y_size
    if cfg.get("absent_rank") is not None:
        # A rank that never makes the matching collective call must not wedge its peers (exchange.cuh): the others give
        # up after LAPS_XCHG_TIMEOUT_S with an error through laps_last_error, and their handles stay dead.
        from laps_b200 import capi
        if rank != cfg["absent_rank"]:
            for _ in range(2):
                try:
                    g.set_primitive(prim[:, zo:zo + zn])
                    raise SystemExit("the collective returned although a rank was absent")
                except capi.LapsError as e:
                    assert "slab exchange aborted" in str(e), str(e)
        dist.barrier()
        g.close()
        dist.barrier()
        dist.destroy_process_group()
        print(f"rank {rank}/{world} ok")
        return
    rows = g.ky_rows                      # this rank's Fourier rows (a contiguous slab unless LAPS_TUNE_CYCLIC=1)
    cyclic = g.ext.y_stride > 1
    if "expect_stride" in cfg:
        assert g.ext.y_stride == cfg["expect_stride"], (g.ext.y_stride, cfg["expect_stride"])

    # decomposition tables: decompose_1d gives n/P each, the remainder to the last rank (parallel.f90:326-349)
    q = nz // world
    assert zo == rank * q and zn == (q if rank < world - 1 else nz - (world - 1) * q)
    q = ny // world
    if cyclic:
        assert np.array_equal(rows, np.arange(rank, ny, world))
    else:
        assert yo == rank * q and yn == (q if rank < world - 1 else ny - (world - 1) * q)

    # transpose_yz index map against SURVEY 9.9: element (gx, gy, gz) of w_yxz goes to the owner of gy,
    # local offset gx + nxh*((gy - yj_off) + yj_size*gz) in the reference; this library keeps z fastest:
    # (gx*yj_size + (gy - yj_off))*nz + gz.
    m = g.transpose_yz_indexmap().reshape(g.nxh, ny, zn, 2)
    offs, sizes = lo.decompose_1d(ny, world)
    owner = np.minimum(np.arange(ny) // (ny // world), world - 1)
    kx, ky, zl = np.meshgrid(np.arange(g.nxh), np.arange(ny), np.arange(zn), indexing="ij")
    if cyclic:   # the experimental round-robin ownership: ky -> rank ky % P, local row ky // P
        cnt = np.array([len(range(r, ny, world)) for r in range(world)])
        assert np.array_equal(m[..., 0], ky % world)
        assert np.array_equal(m[..., 1], (kx * cnt[ky % world] + ky // world) * nz + zo + zl)
    else:
        assert np.array_equal(m[..., 0], owner[ky])
        assert np.array_equal(m[..., 1], (kx * np.asarray(sizes)[owner[ky]] + (ky - np.asarray(offs)[owner[ky]])) * nz + zo + zl)

    # transpose_zy (parallel.f90:300-324), the inverse direction: element (gx, gy, gz) of w_zxy goes to the owner of gz,
    # local offset gx + nxh*(gy + ny*(gz - zj_off)) in the reference's w_yxz; this library keeps z fastest:
    # (gx*ny + gy)*zj_size + (gz - zj_off).
    m2 = g.transpose_zy_indexmap().reshape(g.nxh, yn, nz, 2)
    zoffs, zsizes = lo.decompose_1d(nz, world)
    zowner = np.minimum(np.arange(nz) // (nz // world), world - 1)
    kx2, kyl2, z2 = np.meshgrid(np.arange(g.nxh), np.arange(yn), np.arange(nz), indexing="ij")
    assert np.array_equal(m2[..., 0], zowner[z2])
    assert np.array_equal(m2[..., 1], (kx2 * ny + rows[kyl2]) * np.asarray(zsizes)[zowner[z2]] + (z2 - np.asarray(zoffs)[zowner[z2]]))

    # FFT of a position-encoding field (the idea of ipert=999, mhdinit.f90:1021-1030)
    rng = np.random.default_rng(7)
    a = rng.standard_normal((2, nz, ny, nx))
    w = g.fft_forward(a[:, zo:zo + zn])
    ref = lo.fft_forward(a)
    assert pc.rel_l2(w, ref[:, :, rows, :]) < 1e-13
    b = g.fft_inverse(np.ascontiguousarray(ref[:, :, rows, :]))
    assert pc.rel_l2(b, a[:, zo:zo + zn]) < 1e-13

    # the RK step, driven like mhd.f90: vardt; nsteps x (evolve; evolve_radius; vardt)
    o = pc.oracle_state(p)
    o.set_primitive(prim)
    g.set_primitive(prim[:, zo:zo + zn])
    o.vardt()
    g.vardt()
    assert abs(g.dt - o.dt) <= 1e-13 * o.dt, (g.dt, o.dt)
    for _ in range(cfg.get("steps", 1)):
        o.step()
        g.step()
    tol = cfg.get("tol", 1e-11)
    uu, prim_g = g.get_state()
    for v in range(8):
        assert pc.rel_l2(uu[v], o.uu[v, zo:zo + zn]) < tol, (v, pc.rel_l2(uu[v], o.uu[v, zo:zo + zn]))
    uf = g.uu_fourier()
    for v in range(8):
        assert pc.rel_l2(uf[v], o.uu_fourier[v][:, rows, :]) < tol * max(1.0, np.linalg.norm(o.uu_fourier[v]) / max(np.linalg.norm(o.uu_fourier[v][:, rows, :]), 1e-300))
    assert abs(g.dt - o.dt) <= 1e-12 * o.dt
    # diagnostics are global (allreduce) and identical on every rank
    ave, rms, ru2 = g.calc_rms()
    oave, orms, oru2 = o.calc_rms()
    assert np.allclose(ave, oave, rtol=1e-9, atol=1e-12) and np.allclose(rms, orms, rtol=1e-9, atol=1e-15)
    assert np.allclose(ru2, oru2, rtol=1e-9, atol=1e-18)
    inv = g.invariants()
    oinv = o.invariants()
    assert abs(inv[0] - oinv[0]) <= 1e-9 * abs(oinv[0])
    assert g.checkNan() is False          # checkNan's MPI_Allreduce(MAX) over the ranks (2D/mhd.f90:563-591)
    extra = []
    if p.incompressible:   # the divergence diagnostics of the incompressible driver (mhd.f90:620-732)
        dv, odv = g.calc_max_divV(), o.calc_max_divV()
        assert abs(dv - odv) <= 1e-9 * odv, (dv, odv)
        dr, odr = g.calc_max_div_real(), o.calc_max_div_real()
        assert abs(dr[1] - odr[1]) <= 1e-9 * odr[1], (dr, odr)
        extra = [dv, dr[0], dr[1], g.rho0]
    t = torch.tensor(list(ave) + list(rms) + list(ru2) + list(inv) + [g.dt] + extra, dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    ts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(ts, t)
    for other in ts:
        assert torch.equal(other, ts[0]), "diagnostics differ between ranks"
    dist.barrier()
    g.close()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
