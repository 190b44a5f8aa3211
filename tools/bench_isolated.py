"""Bench legs that run in their OWN process (bench.py starts this file as a child at N = 1 and merges the JSON it prints).

Why a child: these kernels were written when the round's GPU budget was all but spent (one 3.7 s run of the torch-free
checker, profiles/r02_cull_gpu_check.json, is all they saw of a B200), so this leg's first execution is the driver's own bench run.  A fault in them must not be able to take the headline measurement down with it: the child has
its own CUDA context, a time limit, and whatever it prints (or fails to print) only fills the `mesh_cull` / `render_metrics` keys.

Leg 1: eval_rendering's per-frame metrics (usl_render_metrics) on a full-resolution frame.
Leg 2: mesh culling (src/tools/cull_mesh.py) at Replica size -- a marching-cubes mesh of the synthetic room (1.25 cm grid,
about a million vertices) against 64 full-resolution (1200x680) depth frames of one lap of the trajectory:
usl_mesh_cull_frames with and without the occlusion test, the convex-bound test, the face rule + compaction; the kernel's
marks are compared with the host harness (tests/host_harness: the kernels' own thread functions compiled with g++) on a
20 000-vertex sample.  Timing: CUDA events on the launching stream, after a warm-up call.
"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, out


def leg_mesh_cull(P, dev, n_frames=64, voxel=0.0125, sample=20000):
    syn, meshmod = P.synthetic, P.mesh
    cfg = syn.CONFIGS["replica_room0"]
    cam = cfg.cam
    seq = syn.SyntheticSequence(cfg, n_frames=800, device=dev)
    ks = [int(k) for k in np.linspace(0, 799, n_frames)]
    depths = torch.stack([seq.frame(k)[1] for k in ks]).contiguous()
    c2ws = torch.stack([seq.poses[k] for k in ks]).contiguous().float()
    # the mesh: marching cubes (usl_mc_*) of the analytic SDF on a `voxel` grid over the meshing bound (+- 0.05, Mesher.py:178-193)
    axes = []
    for lo, hi in cfg.bound_yaml:
        n = int(round((hi - lo + 0.1) / voxel))
        axes.append(torch.from_numpy(np.linspace(lo - 0.05, hi + 0.05, n)).float().to(dev))
    nx, ny, nz = [a.numel() for a in axes]
    vol = torch.empty((ny, nx, nz), device=dev, dtype=torch.float32)
    for y0 in range(0, ny, 16):
        y1 = min(y0 + 16, ny)
        gy, gx, gz = torch.meshgrid(axes[1][y0:y1], axes[0], axes[2], indexing="ij")
        vol[y0:y1] = seq.room.sdf(torch.stack([gx, gy, gz], dim=-1))
    verts, faces = meshmod.MeshExtractor(axes).run(vol.contiguous())
    V, T = int(verts.shape[0]), int(faces.shape[0])
    culler = meshmod.MeshCuller(cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy, cfg.truncation)
    res = {"workload": f"marching-cubes mesh of the synthetic room ({voxel * 100:.2f} cm grid: {V} vertices, {T} faces) against {n_frames} "
                       f"depth frames {cam.W}x{cam.H} ({depths.numel() * 4 / 2**20:.0f} MiB resident)", "vertices": V, "faces": T, "frames": n_frames}
    w2c = torch.inverse(c2ws).contiguous()
    from host_harness import loader
    pick = torch.linspace(0, V - 1, min(sample, V), device=dev).long()
    v_host = verts[pick].cpu().numpy()
    depths_host = depths.cpu().numpy()
    for tag, eval_rec in (("occlusion", True), ("frustum_only", False)):
        ms, seen = _timed(lambda: culler.seen_by_frames(verts, c2ws, depths, eval_rec, w2c=w2c))       # zero-fill of the marks + the launch
        ref = loader.cull_host_frames(v_host, w2c.cpu().numpy(), depths_host, (cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy),
                                      cfg.truncation, eval_rec, 16)
        got = seen[pick].cpu().numpy()
        res[f"cull_frames_{tag}_ms"] = ms
        res[f"cull_frames_{tag}_seen_fraction"] = float(seen.float().mean())
        res[f"cull_frames_{tag}_mismatch_vs_host_harness"] = int((got != ref).sum())
        res[f"cull_frames_{tag}_vertex_frame_tests_per_s"] = V * n_frames / (ms * 1e-3)     # upper count: early exits do fewer
    res["harness_sample"] = int(pick.numel())
    # frames per CTA = how many depth frames the CTAs in flight share in L2 (1: every group streams one frame past all vertices)
    sweep = {}
    for fpc in (1, 4, 16, 64):
        culler.frames_per_cta = fpc
        sweep[str(fpc)], _ = _timed(lambda: culler.seen_by_frames(verts, c2ws, depths, True, w2c=w2c), iters=2)
    culler.frames_per_cta = 0
    res["cull_frames_occlusion_ms_by_frames_per_cta"] = sweep
    ms, out = _timed(lambda: culler.cull_by_frames(verts, faces, None, c2ws, depths, True))
    res["cull_mesh_total_ms"] = ms                                       # marks + face rule + 2 scans + compaction + the size read
    res["culled_vertices"], res["culled_faces"] = int(out[0].shape[0]), int(out[1].shape[0])
    # convex bound: the room's box pulled in by 30 cm, as 12 triangles (what a trimesh hull holds)
    lo = np.array([b[0] + 0.3 for b in cfg.bound_yaml]); hi = np.array([b[1] - 0.3 for b in cfg.bound_yaml])
    hv = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    hf = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]])
    planes = meshmod.hull_planes(hv, hf)
    ms, inside = _timed(lambda: culler.inside_hull(verts, planes))
    res["cull_hull_ms"] = ms
    res["cull_hull_planes"] = int(planes.shape[0])
    res["cull_hull_mismatch_vs_host_harness"] = int((inside[pick].cpu().numpy() != loader.cull_host_hull(v_host, planes)).sum())
    inside_box = ((verts >= torch.from_numpy(lo).float().to(dev)) & (verts <= torch.from_numpy(hi).float().to(dev))).all(dim=1)
    res["cull_hull_mismatch_vs_box_test"] = int((inside.bool() != inside_box).sum())     # differs only by rounding on the faces
    return res


def leg_render_metrics(P, dev, frames=8):
    """eval_rendering's per-frame PSNR / depth L1 (usl_render_metrics) on full-resolution frames: synthetic sensor frames
    against a perturbed copy standing in for the rendering; sums compared with the host harness."""
    from host_harness import loader
    syn = P.synthetic
    cfg = syn.CONFIGS["replica_room0"]
    seq = syn.SyntheticSequence(cfg, n_frames=800, device=dev)
    col, dep, _ = seq.frame(0)
    g = torch.Generator(device=dev).manual_seed(9)
    ren_c = (col + 0.05 * torch.randn(col.shape, device=dev, generator=g)).clamp(0, 1).contiguous()
    ren_d = (dep + 0.02 * torch.randn(dep.shape, device=dev, generator=g)).contiguous()
    rm = P.RenderMetrics(dev, max_frames=frames + 1)
    rm.add(ren_c, ren_d, col, dep)                                        # warm-up, and the frame that is checked
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        rm.add(ren_c, ren_d, col, dep)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frames
    ref = loader.metrics_host(col.cpu().numpy(), dep.cpu().numpy(), ren_c.cpu().numpy(), ren_d.cpu().numpy())
    got = rm.acc[0].cpu().numpy()
    r = rm.result()
    n = dep.numel()
    return {"pixels": n, "ms_per_frame": ms, "streamed_gbs": n * 32 / (ms * 1e-3) / 1e9, "pixels_with_depth": int(got[2]),
            "max_rel_diff_vs_host_harness": float(np.max(np.abs(got[:2] - ref[:2]) / np.abs(ref[:2]))), "count_matches": bool(got[2] == ref[2]),
            "psnr": float(r["psnr"][0]), "depth_l1": float(r["depth_l1"][0])}


def main():
    t0 = time.time()
    out = {}
    try:
        P = importlib.import_module("uni-slam_b200")
        torch.cuda.set_device(0)
    except Exception as e:                                    # noqa: BLE001 -- reported, never raised into the parent
        print("ISOLATED_JSON " + json.dumps({"isolated_legs_error": f"{type(e).__name__}: {e}"[:400]}), flush=True)
        return
    for key, leg in (("render_metrics", leg_render_metrics), ("mesh_cull", leg_mesh_cull)):
        t1 = time.time()
        try:
            out[key] = leg(P, "cuda:0")
            out[key]["seconds"] = time.time() - t1
        except Exception as e:                                # noqa: BLE001
            out[key] = {"error": f"{type(e).__name__}: {e}"[:400]}
    out["isolated_legs_seconds"] = time.time() - t0
    print("ISOLATED_JSON " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
