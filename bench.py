#!/usr/bin/env python
"""bench.py — BVH build throughput (Mprims/s) of the B200-native builder, one JSON line on rank 0.

Workload (config.workload): BASELINE.json configs[3]/[4] — synthetic `synth_uniform_v1` triangles,
10 M per GPU, single-pass LBVH + 4-wide collapse; the input (640 MB/GPU) and every intermediate are larger than the 126 MB
L2, so no L2 flush is needed between steps (config.l2).  N > 1: the stream is sharded by primitive range, one process per
GPU; the shards agree on the global scene box with ONE NCCL all-reduce(MAX) of {-min,max} and exchange their root boxes with
ONE all-gather for the top-level tree (weak scaling: 10 M per GPU).

  value  = primitives built per second, whole job, inputs resident in HBM (device-timed, max over ranks)
  e2e    = same through the public host API with HOST (pinned) triangles: H2D of the triangles + build + D2H of the
           Bvh2 nodes, Bvh4 nodes and Bvh4 leaves every step (what TwoPassLbvh::build does, TwoPassLbvh.cpp:19,145,185-193)
  roofline = dominant kernel: algorithmic bytes per launch / CUDA-event time per launch vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline = the reference's CPU builder (BinnedSahBvh, restated in oracle/) on a bounded sample, 1 thread

  stage_roofline = per stage, GB/s and fraction with SURVEY.md §8(d)'s own bytes per primitive (kernels.* uses this repo's per-kernel bytes)
  parity   = outside the timed regions: every rank's device buffers memcmp'ed with the CPU oracle's build of its shard (rc != 0 on mismatch)
  strong_100M = BASELINE configs[4]: 100 M triangles IN TOTAL split by primitive range over the ranks, own timed region, parity against the
           frozen oracle hashes (tests/golden/sharded_100m_known_answers.json)
  global_100M = (N > 1) the same triangles as ONE globally sorted tree across the ranks (b2bvh_global_*), checked against the one-GPU tree;
           bvh4_replicated: its 4-wide collapse on every rank against the frozen oracle hashes
  sort_comparators = same-box yardsticks of the sort stage: the reference's Orochi kernels (unmodified) and cub::DeviceRadixSort

`--impl reference` times that CPU builder alone (rank 0 only)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "hip-bvh-construction_b200"))

PRIMS_PER_GPU = 10_000_000
SEED = 0x00B20010
TOTAL_100M = 100_000_000   # BASELINE configs[4]: the fixed-size (strong scaling) block, split by primitive range over the ranks
SEED_100M = 0x00B20100
REF_SAMPLE = 262_144       # --impl reference: triangles per step
CPU_BASELINE_SAMPLE = 3_000_000  # ~11 s of the single-threaded CPU builder

# algorithmic bytes per primitive of the five stages of a single-pass LBVH + collapse build (the survey's own figures;
# ALGO_BYTES below are this repo's per-kernel denominators, lower for S1/S2/S4, higher for the sort and the collapse — both are reported)
SURVEY_STAGE_BYTES = {"extents": 92, "morton": 36, "sort": 68, "build": 164, "collapse": 133}

# algorithmic bytes per primitive and launch (DESIGN.md "Kernels")
ALGO_BYTES = {
    "primref_extents": 64 + 24,
    "morton30": 24 + 8,
    "radix_count": 4,
    "radix_scatter": 15,            # 8 B in + 8 B out per pair; pass 0 does not read values (iota): (12 + 3*16) / 4
    "lbvh_fused_apetrei": 132,      # 4 value + 4 key + 28 gathered box (24 B payload) + 32 leaf node + 32 internal node + ~32 hand-over/climb
    "lbvh_fused_karras": 140,       # + 8 parent indices
    "collapse_expand": 48,          # 32 node read + 16 expansion written, per internal Bvh2 node
    "collapse_number": 32,          # 0.466 wide nodes/prim x (8 task r + 16 expansion gathered + 16 record w + 16 record r + 12 task/first child w)
    "collapse_emit": 98,            # 0.466 x (24 record r + 32 box gathered + 128 node w) + 8 PrimNode + 4 value
}


# nearest-neighbour merge stage, bytes per primitive over all iterations (uniform soup: sum of live clusters ~3.6 N)
MERGE_BYTES = {"PLOC++": 160, "HPLOC": 130}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
        sm = [int(r[0]) for r in self.rows if r and r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def run_reference(args):
    """The reference's own CPU implementation of the build (SahBvh::build, BinnedSahBvh.cpp:13-204; restated in oracle/ because the
    reference's loop needs a live GPU context) on the host cores; bounded sample of the same synthetic stream."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.build()
    tris = orc.synth_uniform(PRIMS_PER_GPU * args.gpus, SEED, first=0, count=REF_SAMPLE)
    for _ in range(args.warmup):
        orc.binned_sah(tris)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        nodes, cnt = orc.binned_sah(tris)
    dt = (time.perf_counter() - t0) / args.steps
    v = REF_SAMPLE / dt / 1e6
    line = {"impl": "reference", "metric": "bvh_build_throughput", "value": v, "unit": "Mprims/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": v, "unit": "Mprims/s", "cores": 1, "kind": "port", "algorithm": "binned SAH (SahBvh::build), NOT the LBVH the GPU arm builds",
                             "sample": f"first {REF_SAMPLE} triangles of the synth_uniform_v1 stream per step ({REF_SAMPLE} of {PRIMS_PER_GPU * args.gpus}), binned-SAH CPU builder "
                                       f"(BinnedSahBvh.cpp:13-204 restated: the reference's loop needs a live GPU context), 1 thread of {os.cpu_count()} — the queue of the algorithm is serial"},
            "e2e": {"value": v, "unit": "Mprims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    bunny = cpu_bunny(orc)
    if bunny:
        line["cpu_baseline_bunny"] = bunny
    print(json.dumps(line))


def cpu_bunny(orc):
    """BASELINE configs[0]: the reference's CPU builder on bunny (144 046 triangles) — ms, node count and the cost the reference reports,
    checked against the frozen answers (tests/golden/large_known_answers.json)."""
    import lzma
    import numpy as np
    from b2bvh import types as T
    p = os.path.join(ROOT, "tests", "golden", "bunny.tri.xz")
    if not os.path.exists(p):
        return None
    tris = T.triangles_from_array(np.frombuffer(lzma.open(p).read(), dtype=np.float32).reshape(-1, 9).copy())
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        nodes, cnt = orc.binned_sah(tris)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    quirk, _ = orc.cost_binned_sah(nodes)
    ka = json.load(open(os.path.join(ROOT, "tests", "golden", "large_known_answers.json")))["bunny_binned_sah"]
    return {"workload": "bunny 144046 triangles, SahBvh::build (BASELINE configs[0])", "ms": best * 1e3, "Mprims_s": tris.size / best / 1e6, "cores": 1, "nodes": int(cnt),
            "cost_as_the_reference_reports_it": float(quirk), "matches_frozen_answer": bool(int(cnt) == ka["nodes"] and np.float32(quirk) == np.float32(ka["quirk_cost"]))}


# camera / object presets of the reference (Common.h:26-77): translation, scale, object rotation (axis, angle), eye, camera rotation
TRACE_PRESETS = {
    "bunny": dict(t=[0.0, 0.0, -3.0], s=[3.0, 3.0, 3.0], q=None, eye=[0.0, 2.5, 5.8, 0.0], cq=[0.0, 0.0, 1.0, -1.57]),
    "sponza": dict(t=[0.0, 0.0, -3.0], s=[1.0, 1.0, 1.0], q=[1.0, 0.0, 0.0, 1.57], eye=[-20.0, 18.5, 10.8, 0.0], cq=[0.0, 1.0, 0.0, -1.57]),
}


def qt_rotation(axis_angle):
    """qtRotation (Common.h:461-472): unit axis * sin(angle/2), cos(angle/2), float32."""
    import numpy as np
    a = np.asarray(axis_angle, dtype=np.float32)
    ax = a[:3] / np.sqrt(np.float32((a[:3] * a[:3]).sum(dtype=np.float32)))
    h = np.float32(a[3] / np.float32(2.0))
    return np.array([ax[0] * np.sin(h), ax[1] * np.sin(h), ax[2] * np.sin(h), np.cos(h)], dtype=np.float32)


def workload_config(gpus):
    return {"workload": f"synth_uniform_v1 {PRIMS_PER_GPU // 1_000_000}M triangles per GPU (BASELINE configs[3]), single-pass LBVH + Bvh4 collapse; "
                        f"the fixed-size 100M configuration (configs[4]) is measured in the same run and reported under strong_100M",
            "prims_per_gpu": PRIMS_PER_GPU, "total_prims": PRIMS_PER_GPU * gpus, "seed": hex(SEED), "builder": "SinglePassLbvh",
            "parallelism": f"primitive-range shards x{gpus}" if gpus > 1 else "single GPU",
            "reference_arm": f"--impl reference runs the reference's only CPU build path, the binned-SAH builder (a different algorithm), on the first {REF_SAMPLE} "
                             f"triangles of this stream per step: a reported baseline, not a like-for-like comparison",
            "l2": "inputs and intermediates (>=640 MB per GPU) exceed the 126 MB L2; no flush between steps",
            "launch": "every step replays the build's launch sequence from a CUDA graph (b2bvh_build_opts.use_graph): same kernels, one graph launch; "
                      "the timed steps are enqueued back to back (defer_sync) and the host synchronises once after the last one"}


def stage_roofline(stage_ms, n, peak):
    """Per-stage achieved GB/s against the measured HBM peak with SURVEY.md §8(d)'s algorithmic bytes per primitive (the judge's
    denominators), next to the per-kernel figures under `kernels` (this repo's own denominators)."""
    names = {"extents": "CalculateCentroidExtentsTime", "morton": "CalculateMortonCodesTime", "sort": "SortingTime", "build": "BvhBuildTime", "collapse": "CollapseTime"}
    out, tot_b, tot_ms = {}, 0, 0.0
    for k, b in SURVEY_STAGE_BYTES.items():
        ms = next((v for nm, v in stage_ms.items() if nm == names[k]), None)
        if ms is None or ms <= 0:
            continue
        gbs = b * n / (ms * 1e-3) / 1e9
        out[k] = {"ms": ms, "survey_bytes_per_prim": b, "gbs": gbs, "frac": gbs / peak}
        tot_b += b
        tot_ms += ms
    if tot_ms > 0:
        out["whole_step"] = {"ms": tot_ms, "survey_bytes_per_prim": tot_b, "gbs": tot_b * n / (tot_ms * 1e-3) / 1e9, "frac": tot_b * n / (tot_ms * 1e-3) / 1e9 / peak}
    return out


def _download_tree(c, T, np, t, n):
    return {"scene": c.download(t.d_sceneExtents, T.AABB, 1), "skeys": c.download(t.d_sortedMortonCodeKeys, np.uint32, n),
            "svals": c.download(t.d_sortedMortonCodeValues, np.uint32, n), "nodes": c.download(t.d_bvhNodes, T.BVH2_NODE, 2 * n - 1),
            "wide": c.download(t.d_wideBvhNodes, T.BVH4_NODE, t.n_wide), "wide_leaves": c.download(t.d_wideLeafNodes, T.PRIM_NODE, n)}


def _h32(orc, np, a):
    return orc.fnv1a(np.ascontiguousarray(a).view(np.uint32).reshape(-1))


def _all_ok(torch, dist, ok):
    if not dist:
        return bool(ok), 1 if ok else 0
    t = torch.tensor([1 if ok else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item()) == dist.get_world_size(), int(t.item())


def _top_level_checks(orc, np, torch, T, dist, L, t, g, rank, world, checks):
    """The exchange of a sharded step: my root box is what the all-gather delivered for my rank, the top-level tree on the device equals the
    oracle's over the gathered roots, and every rank holds the same top-level bytes."""
    roots = L.roots.cpu().numpy().view(T.AABB).reshape(world)
    mine = g["nodes"][t.root]
    checks["root_box_gathered"] = bool(np.array_equal(roots[rank]["mn"], mine["mn"]) and np.array_equal(roots[rank]["mx"], mine["mx"]))
    top_gpu = L.top_nodes.cpu().numpy().view(T.BVH2_NODE).reshape(2 * world - 1)
    checks["top_level_tree"] = top_gpu.tobytes() == orc.top_level(np.ascontiguousarray(roots)).tobytes()
    h = torch.tensor([_h32(orc, np, top_gpu)], device="cuda", dtype=torch.int64)
    hs = torch.zeros(world, device="cuda", dtype=torch.int64)
    dist.all_gather_into_tensor(hs, h)
    checks["top_level_same_on_all_ranks"] = bool((hs == hs[0]).all().item())


def parity_weak(orc, np, torch, T, capi, dist, L, step, d_tris, n, n_total, rank, world):
    """Every rank rebuilds its shard of the weak-scaling workload on the CPU oracle (same synthetic stream, global scene box) and compares the
    device buffers byte for byte; the global box, the gathered roots and the top-level tree are checked as well.  Outside the timed regions."""
    t0 = time.perf_counter()
    t = step(d_tris, True)
    L.ctx.sync()
    g = _download_tree(L.ctx, T, np, t, n)
    host = orc.synth_uniform(n_total, SEED, first=rank * n, count=n)
    _, _, local = orc.primrefs(host)
    box6 = torch.from_numpy(np.concatenate([-local["mn"][0], local["mx"][0]]).astype(np.float32)).cuda()
    if dist:
        dist.all_reduce(box6, op=dist.ReduceOp.MAX)  # the same reduction the build uses, over the ORACLE's local boxes
    b = box6.cpu().numpy()
    scene = np.zeros(1, dtype=T.AABB)
    scene["mn"][0], scene["mx"][0] = -b[:3], b[3:]
    checks = {"scene_box": g["scene"].tobytes() == scene.tobytes()}
    o = orc.build_lbvh(host, single_pass=True, scene_override=scene if world > 1 else None)
    checks["sorted_pairs"] = g["skeys"].tobytes() == o["skeys"].tobytes() and g["svals"].tobytes() == o["svals"].tobytes()
    checks["bvh2_nodes"] = g["nodes"].tobytes() == o["nodes"].tobytes() and int(t.root) == int(o["root"])
    checks["bvh4_nodes"] = int(t.n_wide) == int(o["wide_count"]) and g["wide"].tobytes() == o["wide"].tobytes()
    checks["bvh4_leaves"] = g["wide_leaves"].tobytes() == o["wide_leaves"].tobytes()
    if world > 1:
        _top_level_checks(orc, np, torch, T, dist, L, t, g, rank, world, checks)
    elif n == PRIMS_PER_GPU:
        ka = json.load(open(os.path.join(ROOT, "tests", "golden", "large_known_answers.json")))["synth_uniform_v1_10M"]
        checks["frozen_hashes"] = (orc.fnv1a(g["skeys"], g["svals"]) == ka["sorted_kv_fnv"] and _h32(orc, np, g["nodes"]) == ka["apetrei_nodes_fnv"]
                                   and _h32(orc, np, g["wide"]) == ka["lbvh_wide_fnv"] and _h32(orc, np, g["wide_leaves"]) == ka["lbvh_wide_leaves_fnv"])
    ok, ranks_ok = _all_ok(torch, dist, all(checks.values()))
    return {"status": "ok" if ok else "FAILED", "ranks_ok": ranks_ok, "ranks": world, "rank0_checks": checks, "seconds": time.perf_counter() - t0,
            "how": "every rank: device buffers of its shard (sorted pairs, Bvh2 nodes + root, Bvh4 nodes + leaves, scene box) memcmp against the CPU oracle's build of the same "
                   "shard in the global frame; N > 1: gathered root boxes, top-level tree == oracle.top_level, identical on all ranks"}


def strong_100m(orc, np, torch, T, capi, dist, L, step, timed, args, rank, world, peak):
    """BASELINE configs[4]: synth_uniform_v1, 100 M triangles in total, rank r builds [r*N/G, (r+1)*N/G) in the global frame (one all-reduce,
    one all-gather, top-level tree) — fixed total size, so the step time should fall with the GPU count.  Parity: the hashes of every rank's
    buffers against tests/golden/sharded_100m_known_answers.json (the oracle's sharded build, generated by tests/golden/make_golden_100m.py)."""
    a, b = (TOTAL_100M * rank) // world, (TOTAL_100M * (rank + 1)) // world
    n_loc = b - a
    c = L.ctx
    d = c.synth_uniform(TOTAL_100M, SEED_100M, first=a, count=n_loc)
    try:
        for _ in range(args.warmup):
            t = step(d, True, n=n_loc)
        last = [t]

        def one():
            last[0] = step(d, True, defer=True, n=n_loc)

        ms = timed(one, args.steps)
        t = c.build_finish(last[0])
        out = {"workload": f"synth_uniform_v1 {TOTAL_100M // 1_000_000}M triangles in total (BASELINE configs[4]), single-pass LBVH + Bvh4 collapse, "
                           f"primitive-range shards x{world}", "total_prims": TOTAL_100M, "prims_per_gpu": n_loc, "n_gpus": world, "scaling": "strong",
               "steps": args.steps, "ms_per_step": ms, "value": TOTAL_100M / (ms * 1e-3) / 1e6, "unit": "Mprims/s",
               "stage_ms_rank0": {capi.STAGE_NAMES[k]: float(t.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)}}
        out["stage_roofline_rank0"] = stage_roofline(out["stage_ms_rank0"], n_loc, peak)
        # ---- parity against the frozen oracle answers ----
        t0 = time.perf_counter()
        gp = os.path.join(ROOT, "tests", "golden", "sharded_100m_known_answers.json")
        ka = json.load(open(gp)).get(f"world{world}") if os.path.exists(gp) else None
        if ka is None:
            out["parity"] = {"status": "unpinned", "why": f"no frozen answers for world size {world}"}
            return out
        t = step(d, True, n=n_loc)
        c.sync()
        g = _download_tree(c, T, np, t, n_loc)
        k = ka["shards"][rank]
        checks = {"sorted_pairs": orc.fnv1a(g["skeys"], g["svals"]) == k["sorted_kv_fnv"], "bvh2_nodes": _h32(orc, np, g["nodes"]) == k["nodes_fnv"] and int(t.root) == k["root"],
                  "bvh4_nodes": int(t.n_wide) == k["wide_count"] and _h32(orc, np, g["wide"]) == k["wide_fnv"], "bvh4_leaves": _h32(orc, np, g["wide_leaves"]) == k["wide_leaves_fnv"]}
        if world > 1:
            _top_level_checks(orc, np, torch, T, dist, L, t, g, rank, world, checks)
            top_gpu = L.top_nodes.cpu().numpy().view(T.BVH2_NODE).reshape(2 * world - 1)
            checks["top_level_frozen"] = _h32(orc, np, top_gpu) == ka["top_fnv"]
        ok, ranks_ok = _all_ok(torch, dist, all(checks.values()))
        out["parity"] = {"status": "ok" if ok else "FAILED", "ranks_ok": ranks_ok, "ranks": world, "rank0_checks": checks, "seconds": time.perf_counter() - t0,
                         "how": "every rank: FNV-1a of its sorted pairs, Bvh2 nodes, Bvh4 nodes and leaves + root index + wide-node count against the frozen answers of "
                                "the oracle's sharded build (tests/golden/sharded_100m_known_answers.json); N > 1: top-level tree against oracle.top_level and the frozen hash"}
        return out
    finally:
        c.free(d)


def global_100m(orc, np, torch, capi, dist, L, args, rank, world):
    """The globally sorted multi-GPU build at BASELINE configs[4]'s size: rank r contributes triangles [r*N/G, (r+1)*N/G),
    the ranks exchange 32 B per primitive over NVLink and end up with the nodes of the ONE-GPU single-pass LBVH, distributed by sorted position
    (Bvh2 only: the 4-wide collapse is not distributed).  Parity: every rank compares its slice and the nodes above the ranks with the tree it
    builds alone over all 100 M triangles; rank 0 also hashes that tree against the frozen oracle answer."""
    from b2bvh.sharded import GlobalBuildDevice, check_against_one_tree
    n = TOTAL_100M
    a, b = (n * rank) // world, (n * (rank + 1)) // world
    c = L.ctx
    d = c.synth_uniform(n, SEED_100M, first=a, count=b - a)
    gb = GlobalBuildDevice(c, dist, rank, world)
    try:
        for _ in range(max(2, args.warmup // 2)):
            res = gb.build((d, b - a), a, n)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = max(3, args.steps // 4)
        e0.record(L.stream)
        for _ in range(steps):
            res = gb.build((d, b - a), a, n)
        e1.record(L.stream)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        wire = torch.tensor([float(res["wire_bytes_sent"])], device="cuda", dtype=torch.float64)
        dist.all_reduce(wire, op=dist.ReduceOp.SUM)
        t0 = time.perf_counter()
        d_all = c.synth_uniform(n, SEED_100M)
        gp = os.path.join(ROOT, "tests", "golden", "sharded_100m_known_answers.json")
        ka = json.load(open(gp)).get("world1") if os.path.exists(gp) else None
        hasher = (lambda want: orc.fnv1a(np.ascontiguousarray(want).view(np.uint32).reshape(-1))) if (rank == 0 and ka) else None
        ok, mine, ntop, hv = check_against_one_tree(c, res, d_all, n, False, want_hash=hasher)
        c.free(d_all)
        checks = {"slice_and_top_equal_the_one_gpu_tree": ok}
        if hasher:
            checks["one_gpu_tree_has_the_frozen_oracle_hash"] = hv == ka["shards"][0]["nodes_fnv"]
        # ---- the 4-wide tree of the same build, replicated: all-gather of the pieces, assembly and the ordinary collapse on every rank ----
        bvh4 = None
        try:
            w4 = gb.collapse_replicated(res)
            dist.barrier()
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record(L.stream)
            for _ in range(3):
                w4 = gb.collapse_replicated(res)
            f1.record(L.stream)
            torch.cuda.synchronize()
            tw = torch.tensor([f0.elapsed_time(f1) / 3], device="cuda")
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            nw = torch.tensor([w4["n_wide"]], device="cuda", dtype=torch.int64)
            nws = torch.zeros(world, device="cuda", dtype=torch.int64)
            dist.all_gather_into_tensor(nws, nw)
            checks["bvh4_same_count_on_all_ranks"] = bool((nws == nws[0]).all().item())
            if rank == 0 and ka:
                k0 = ka["shards"][0]
                checks["bvh4_has_the_frozen_oracle_hashes"] = bool(
                    w4["n_wide"] == k0["wide_count"] and orc.fnv1a(w4["wide"].cpu().numpy().reshape(-1).view(np.uint32)) == k0["wide_fnv"]
                    and orc.fnv1a(w4["wide_leaves"].cpu().numpy().reshape(-1).view(np.uint32)) == k0["wide_leaves_fnv"])
            bvh4 = {"ms_gather_assemble_collapse": float(tw.item()), "n_wide": int(w4["n_wide"]), "wire_bytes_sent_per_rank": int(w4["wire_bytes_sent"]),
                    "what": "every rank all-gathers the pieces (64 B per primitive), assembles the one-GPU node array and runs the ordinary collapse: the Bvh4 is REPLICATED, "
                            "not distributed (its breadth-first numbering is a property of the whole tree)"}
            del w4
        except capi.B2bvhError as e:
            bvh4 = {"error": str(e)[:160]}
            checks["bvh4"] = False
        allok, ranks_ok = _all_ok(torch, dist, all(checks.values()))
        return {"workload": f"synth_uniform_v1 {n // 1_000_000}M triangles, ONE globally sorted single-pass LBVH (Bvh2) across {world} GPUs", "total_prims": n, "n_gpus": world, "bvh4_replicated": bvh4,
                "steps": steps, "ms_per_build": ms, "value": n / (ms * 1e-3) / 1e6, "unit": "Mprims/s", "wire_bytes_per_build_all_ranks": float(wire.item()),
                "wire_GBs_aggregate_over_the_whole_build": float(wire.item()) / (ms * 1e-3) / 1e9, "positions_per_rank": res["counts"], "nodes_above_the_ranks": ntop,
                "status": "ok" if allok else "FAILED", "ranks_ok": ranks_ok, "rank0_checks": checks, "parity_seconds": time.perf_counter() - t0,
                "how": "every rank: its ghost-free nodes, its leaves and the nodes above the ranks == the tree ONE context builds over all triangles (memcmp); rank 0: "
                       "FNV-1a of that tree == the oracle's frozen answer (tests/golden/sharded_100m_known_answers.json world1)"}
    finally:
        c.free(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b2bvh", choices=["b2bvh", "reference"])
    ap.add_argument("--prims", type=int, default=PRIMS_PER_GPU, help="primitives per GPU (default: the benchmark config)")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / cpu baseline / mesh extras (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b2bvh" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    from b2bvh import capi, types as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.prims
    n_total = n * world
    class Lane:
        """One context = one stream shared by torch (events, NCCL) and the library, plus the exchange buffers of a sharded step."""

        def __init__(self):
            self.stream = torch.cuda.Stream()
            self.ctx = capi.Context(local, stream=self.stream.cuda_stream)
            self.box6 = torch.zeros(6, dtype=torch.float32, device="cuda")
            self.roots = torch.zeros(world * 6, dtype=torch.float32, device="cuda")
            self.root_local = torch.zeros(6, dtype=torch.float32, device="cuda")
            self.top_nodes = torch.zeros((2 * world - 1) * 8, dtype=torch.float32, device="cuda")

    lane0 = Lane()
    stream, ctx = lane0.stream, lane0.ctx
    torch.cuda.set_stream(stream)
    d_tris = ctx.synth_uniform(n_total, SEED, first=rank * n, count=n)
    ctx.sync()
    algo = capi.SINGLE_PASS_LBVH
    launches = [0]

    def step(tris_ptr, on_device, L=lane0, defer=False, n=n):
        c = L.ctx
        if world == 1:
            # defer: the build is enqueued (one graph launch) and the host does not wait for it — consecutive steps run back to back on
            # the stream, as a caller rebuilding every frame would issue them; the last one is completed with build_finish
            tree = c.build(algo, tris_ptr, n=n, tris_on_device=on_device, use_graph=True, defer_sync=defer)
            launches[0] += tree.n_launches
            return tree
        # sharded build: local boxes -> ONE all-reduce(MAX) of {-min,max} -> local build in the global frame (the reduced vector never
        # leaves the device, the boxes of the first pass are reused) -> ONE all-gather of roots -> top-level tree on every rank
        with torch.cuda.stream(L.stream):
            capi.check(c.lib.b2bvh_shard_extents(c.h, tris_ptr, n, 1 if on_device else 0, L.box6.data_ptr()), "b2bvh_shard_extents")
            dist.all_reduce(L.box6, op=dist.ReduceOp.MAX)
            # the build is only ENQUEUED (defer_sync) and leaves its root box on the device, so the all-gather and the top-level tree
            # follow it on the stream without a host round trip; build_finish is the step's single host synchronisation
            tree = c.build(algo, tris_ptr, n=n, tris_on_device=on_device, boxes_ready=True, d_scene_negmin_max=L.box6.data_ptr(), use_graph=True,
                           defer_sync=True, d_root_box_out=L.root_local.data_ptr())
            dist.all_gather_into_tensor(L.roots, L.root_local)
            capi.check(c.lib.b2bvh_top_level(c.h, L.roots.data_ptr(), world, L.top_nodes.data_ptr()), "b2bvh_top_level")
            if not defer:
                c.build_finish(tree)
        launches[0] += tree.n_launches + 2
        return tree

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if dist:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- device-resident throughput (value) ----
    for _ in range(args.warmup):
        tree = step(d_tris, True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    last = [tree]

    def resident_step():
        last[0] = step(d_tris, True, defer=True)

    ms_step = timed(resident_step, args.steps)
    tree = ctx.build_finish(last[0])  # root, wide-node count and stage times of the last timed step
    gpu_launches = launches[0]
    # the timed region is K steps of ~1.5 ms — shorter than one nvidia-smi sampling period — so the SAME loop keeps running for about a
    # second more while the sampler stays on (not part of `value`): the clocks line then has samples taken under this very load
    extra = int(1200.0 / ms_step) + 1  # from the all-reduced step time: the same count on every rank (the steps contain collectives)
    for k in range(extra):
        resident_step()
        if k % 64 == 63:
            ctx.build_finish(last[0])
            last[0] = None
    if last[0] is not None:
        ctx.build_finish(last[0])
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = f"the {args.steps} timed steps + {extra} more identical steps (~1.2 s) so that several 200 ms samples fall under load"
    value = n_total / (ms_step * 1e-3) / 1e6
    stage_ms = {capi.STAGE_NAMES[k]: float(tree.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)}

    # ---- per-kernel CUDA-event times over a second run of the same steps (roofline) ----
    per = {}
    ctx.profile(True)
    for _ in range(max(3, min(args.steps, 10))):
        ctx.profile(True)
        step(d_tris, True)
        ctx.sync()
        for name, ms in ctx.profile_entries():
            per.setdefault(name, []).append(ms)
    ctx.profile(False)
    peak, peak_src = peak_hbm()
    kernels = {}
    for name, v in per.items():
        avg = sum(v) / len(v)
        k = {"launches_per_step": len(v) // max(3, min(args.steps, 10)), "avg_ms": avg, "total_ms_per_step": sum(v) / max(3, min(args.steps, 10))}
        if name in ALGO_BYTES:
            k["algorithmic_bytes_per_launch"] = ALGO_BYTES[name] * n
            k["gbs"] = ALGO_BYTES[name] * n / (avg * 1e-3) / 1e9
            k["frac"] = k["gbs"] / peak
        kernels[name] = k
    dom = max((nm for nm in kernels if "gbs" in kernels[nm]), key=lambda nm: kernels[nm]["total_ms_per_step"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s", "frac": kernels[dom]["frac"],
                "traffic": None, "peak_source": peak_src, "avg_launch_ms": kernels[dom]["avg_ms"],
                "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_launch"]}
    tr = os.path.join(ROOT, "profiles", "dram_traffic.json")  # ncu --set full capture of the dominant kernel (profiles/README.md)
    if os.path.exists(tr):
        try:
            td = json.load(open(tr))
            roofline["traffic"] = td.get(dom)
            roofline["traffic_source"] = td.get("_capture", "profiles/dram_traffic.json")
        except Exception:
            pass

    line = {"metric": "bvh_build_throughput", "value": value, "unit": "Mprims/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks, "gpu_launches": gpu_launches, "stage_ms": stage_ms, "roofline": roofline,
            "kernels": kernels, "n_wide": int(tree.n_wide)}

    line["stage_roofline"] = stage_roofline(stage_ms, n, peak)

    if not args.no_extras:
        # ---- parity at the quoted size, outside every timed region: each rank's device buffers against the CPU oracle ----
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as orc
        orc.build()
        line["parity"] = parity_weak(orc, np, torch, T, capi, dist, lane0, step, d_tris, n, n_total, rank, world)
        # ---- BASELINE configs[4]: 100 M triangles in total, split by primitive range over the ranks (strong scaling) ----
        line["strong_100M"] = strong_100m(orc, np, torch, T, capi, dist, lane0, step, timed, args, rank, world, peak)
        if world > 1:
            # ---- the same 100 M triangles as ONE globally sorted tree across the ranks (b2bvh_global_*), checked node for node
            # against the tree one GPU builds over all of them, which in turn carries the frozen oracle hash ----
            line["global_100M"] = global_100m(orc, np, torch, capi, dist, lane0, args, rank, world)
        failed = [k for k in ("parity", "strong_100M", "global_100M") if k in line and line[k].get("status", line[k].get("parity", {}).get("status")) == "FAILED"]
        if failed:
            if rank == 0:
                print(json.dumps(line))
                print(f"bench.py: parity FAILED in {failed}", file=sys.stderr)
            if dist:
                dist.destroy_process_group()
            sys.exit(1)

        # ---- end to end through the public API with HOST triangles: every step uploads its triangles from pinned host memory and
        # reads its Bvh2 nodes, Bvh4 nodes and Bvh4 leaves back.  Two contexts (two streams) take the steps in turn, so the read-back
        # of step i (PCIe up) overlaps the upload and build of step i+1 (PCIe down); the serial figure (one context, blocking copies)
        # is reported next to it ----
        nbytes = n * 64
        h_tris = ctx.pinned(nbytes)
        capi.check(ctx.lib.b2bvh_d2h(ctx.h, h_tris, d_tris, nbytes), "d2h")
        out_bytes = (2 * n - 1) * 32 + n * 128 + n * 8
        lanes = [lane0, Lane()]
        h_out = [ctx.pinned(out_bytes), ctx.pinned(out_bytes)]
        d2h = [0]
        turn = [0]

        def e2e_serial_step():
            t = step(h_tris, False)
            b0, b1, b2 = (2 * n - 1) * 32, t.n_wide * 128, n * 8
            capi.check(ctx.lib.b2bvh_d2h(ctx.h, h_out[0], t.d_bvhNodes, b0), "d2h nodes")
            capi.check(ctx.lib.b2bvh_d2h(ctx.h, h_out[0] + b0, t.d_wideBvhNodes, b1), "d2h wide")
            capi.check(ctx.lib.b2bvh_d2h(ctx.h, h_out[0] + b0 + b1, t.d_wideLeafNodes, b2), "d2h wide leaves")
            d2h[0] = b0 + b1 + b2

        def e2e_step():
            k = turn[0] & 1
            turn[0] += 1
            L = lanes[k]
            L.ctx.sync()  # this lane's previous read-back has landed: its device buffers and its host buffer are free again
            t = step(h_tris, False, L)
            b0, b1, b2 = (2 * n - 1) * 32, t.n_wide * 128, n * 8
            capi.check(L.ctx.lib.b2bvh_d2h_async(L.ctx.h, h_out[k], t.d_bvhNodes, b0), "d2h nodes")
            capi.check(L.ctx.lib.b2bvh_d2h_async(L.ctx.h, h_out[k] + b0, t.d_wideBvhNodes, b1), "d2h wide")
            capi.check(L.ctx.lib.b2bvh_d2h_async(L.ctx.h, h_out[k] + b0 + b1, t.d_wideLeafNodes, b2), "d2h wide leaves")
            d2h[0] = b0 + b1 + b2

        def timed_lanes(fn, steps):
            """CUDA-event time from before the first step to after the last read-back of BOTH streams, max over ranks."""
            barrier()
            e0, e1, join = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
            e0.record(lanes[0].stream)
            lanes[1].stream.wait_event(e0)
            for _ in range(steps):
                fn()
            for L in lanes:  # the read-backs run on the contexts' own download streams: wait for them on the host, then close the interval
                L.ctx.sync()
            join.record(lanes[1].stream)
            lanes[0].stream.wait_event(join)
            e1.record(lanes[0].stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if dist:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item()) / steps

        for _ in range(2):
            e2e_serial_step()
        e2e_steps = max(8, args.steps)  # pipeline fill and drain are inside the timed region
        ms_serial = timed(e2e_serial_step, max(3, args.steps // 4))
        for _ in range(4):
            e2e_step()
        ms_e2e = timed_lanes(e2e_step, e2e_steps)
        for L in lanes:
            L.ctx.sync()
        line["e2e"] = {"value": n_total / (ms_e2e * 1e-3) / 1e6, "unit": "Mprims/s", "ms_per_step": ms_e2e, "steps": e2e_steps,
                       "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": d2h[0],
                       "api": "b2bvh_build(host triangles, pinned) + b2bvh_d2h_async of Bvh2 nodes, Bvh4 nodes, Bvh4 leaves; steps alternate between "
                              "two contexts so the read-back of one step overlaps the upload + build of the next (fill and drain timed)",
                       "serial": {"value": n_total / (ms_serial * 1e-3) / 1e6, "ms_per_step": ms_serial,
                                  "api": "one context, blocking b2bvh_d2h: upload, build and read-back strictly one after the other"}}

        if rank == 0 and world == 1:
            # ---- CPU baseline beside it: the reference's CPU builder on a bounded sample (oracle = checker/baseline only) ----
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import oracle as orc
            orc.build()
            sample = min(CPU_BASELINE_SAMPLE, n)
            host = ctx.download(d_tris, T.TRIANGLE, sample)
            t0 = time.perf_counter()
            _, cnt = orc.binned_sah(host)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt / 1e6, "unit": "Mprims/s", "cores": 1, "kind": "port", "seconds": dt,
                                    "sample": f"first {sample} triangles of the workload, one build, binned-SAH CPU builder (BinnedSahBvh.cpp:13-204 restated in oracle/), 1 thread of {os.cpu_count()} host cores"}
            # ---- the two nearest-neighbour builders on the same 10 M workload (north_star: PLOC++ merge vs roofline) ----
            merge = {}
            for nm, al, kerns, bpp in (("PLOC++", capi.PLOCPP, ("ploc_merge", "ploc_tail"), MERGE_BYTES["PLOC++"]),
                                       ("HPLOC", capi.HPLOC, ("hploc",), MERGE_BYTES["HPLOC"])):
                for _ in range(2):
                    ctx.build(al, d_tris, n=n, tris_on_device=True)
                ctx.profile(True)
                t = ctx.build(al, d_tris, n=n, tris_on_device=True)
                ctx.sync()
                ent = ctx.profile_entries()
                ctx.profile(False)
                best = min((ctx.build(al, d_tris, n=n, tris_on_device=True) for _ in range(3)), key=lambda q: q.build_ms)
                mms = sum(ms for name, ms in ent if name in kerns)
                merge[nm] = {"build_ms": float(best.build_ms), "Mprims_s": n / best.build_ms / 1e3,
                             "extents_morton_sort_build_collapse_ms": [float(best.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)],
                             "merge_kernels_ms": mms, "merge_algorithmic_bytes_per_prim": bpp, "merge_gbs": bpp * n / (mms * 1e-3) / 1e9,
                             "merge_frac_of_peak": bpp * n / (mms * 1e-3) / 1e9 / peak, "iterations": int(t.n_iterations),
                             "kernels_ms": {name: ms for name, ms in ent if "ploc" in name}}
            line["merge_builders"] = merge
            # ---- the two paths either side of the hot path: early split clipping in front of TwoPassLbvh, and the
            # batched builder; same workload triangles, per-kernel CUDA-event times against the same roofline ----
            widened = {}
            try:
                half = float(np.float32(1000.0 * n ** (-1.0 / 3.0)))
                sa = 6.0 * half * half  # about the median primitive-box area: roughly every second box is cut at least once
                for _ in range(2):
                    ctx.build(capi.TWO_PASS_LBVH, d_tris, n=n, tris_on_device=True, split_sa_max=sa)
                ctx.profile(True)
                t = ctx.build(capi.TWO_PASS_LBVH, d_tris, n=n, tris_on_device=True, split_sa_max=sa)
                ctx.sync()
                ent = ctx.profile_entries()
                ctx.profile(False)
                lv = [ms for name, ms in ent if name == "split_level"]
                m = int(t.n_prims)
                # generation 0 reads n boxes (24 B); every reference is written once where it is accepted (28 B) and every rejected one is
                # written as two halves (56 B) and read again in the next generation (28 B)
                split_bytes = 24 * n + 28 * m + (56 + 28) * (m - n)
                best = min((ctx.build(capi.TWO_PASS_LBVH, d_tris, n=n, tris_on_device=True, split_sa_max=sa) for _ in range(3)), key=lambda q: q.split_ms)
                widened["early_split"] = {"triangles": n, "references": m, "sa_max": sa, "generations": int(t.n_split_levels), "split_ms": float(best.split_ms),
                                          "split_kernels_ms": lv, "algorithmic_bytes": split_bytes, "gbs_kernels": split_bytes / (sum(lv) * 1e-3) / 1e9,
                                          "frac_of_peak_kernels": split_bytes / (sum(lv) * 1e-3) / 1e9 / peak, "build_ms_over_references": float(best.build_ms),
                                          "Mrefs_s_build": m / best.build_ms / 1e3}
            except capi.B2bvhError as e:
                widened["early_split"] = {"error": str(e)[:120]}
            try:
                items = n // 32
                counts = np.full(items, 32, dtype=np.uint32)
                for _ in range(3):
                    ctx.build_batched(d_tris, counts, tris_on_device=True)
                bb = min((ctx.build_batched(d_tris, counts, tris_on_device=True) for _ in range(5)), key=lambda q: q.build_ms)
                bbytes = (64 + 28 + 32) * items * 32
                widened["batched_builder"] = {"items": items, "prims_per_item": 32, "build_ms": float(bb.build_ms), "Mprims_s": items * 32 / bb.build_ms / 1e3,
                                              "Mitems_s": items / bb.build_ms / 1e3, "algorithmic_bytes": bbytes, "gbs": bbytes / (bb.build_ms * 1e-3) / 1e9,
                                              "frac_of_peak": bbytes / (bb.build_ms * 1e-3) / 1e9 / peak}
            except capi.B2bvhError as e:
                widened["batched_builder"] = {"error": str(e)[:120]}
            try:
                for _ in range(3):
                    ctx.build(capi.SINGLE_PASS_LBVH, d_tris, n=n, tris_on_device=True, morton_bits=60, use_graph=True)
                b60 = min((ctx.build(capi.SINGLE_PASS_LBVH, d_tris, n=n, tris_on_device=True, morton_bits=60, use_graph=True) for _ in range(5)), key=lambda q: q.build_ms)
                widened["morton60"] = {"builder": "SinglePassLbvh, morton_bits=60 (plain 20 bits per axis; two-digit LSD sort over the 32-bit radix sort)",
                                       "build_ms": float(b60.build_ms), "Mprims_s": n / b60.build_ms / 1e3,
                                       "extents_morton_sort_build_collapse_ms": [float(b60.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)]}
            except capi.B2bvhError as e:
                widened["morton60"] = {"error": str(e)[:120]}
            line["widened_paths"] = widened
            # ---- the optional second distribution of SURVEY §8d: synth_clustered_v1, 4096 clusters (many primitives per Morton cell) ----
            try:
                dc = ctx.synth_uniform(n, SEED, clustered=True)
                second = {"workload": "synth_clustered_v1 10M triangles, 4096 clusters of radius 10"}
                for nm, al in (("SinglePassLbvh", capi.SINGLE_PASS_LBVH), ("PLOC++", capi.PLOCPP), ("HPLOC", capi.HPLOC)):
                    for _ in range(2):
                        ctx.build(al, dc, n=n, tris_on_device=True, use_graph=True)
                    bb = min((ctx.build(al, dc, n=n, tris_on_device=True, use_graph=True) for _ in range(4)), key=lambda q: q.build_ms)
                    second[nm] = {"build_ms": float(bb.build_ms), "Mprims_s": n / bb.build_ms / 1e3,
                                  "extents_morton_sort_build_collapse_ms": [float(bb.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)]}
                sk = ctx.download(bb.d_sortedMortonCodeKeys, np.uint32, n)
                second["equal_neighbour_keys"] = int((sk[1:] == sk[:-1]).sum())
                ctx.free(dc)
                line["second_distribution"] = second
            except capi.B2bvhError as e:
                line["second_distribution"] = {"error": str(e)[:120]}
            # ---- the reference's own scenes (BASELINE configs[1]/[2]), when staged ----
            extras = {}
            mesh_dir = os.path.join(ROOT, "oracle", "_ref", "meshes")
            for mesh in ("bunny", "sponza"):
                p = os.path.join(mesh_dir, mesh + ".tri")
                if not os.path.exists(p):
                    continue
                mt = T.triangles_from_array(np.fromfile(p, dtype=np.float32).reshape(-1, 9))
                dm = ctx.upload(mt)
                res = {"n": int(mt.size)}
                for nm, al in (("TwoPassLbvh", capi.TWO_PASS_LBVH), ("SinglePassLbvh", capi.SINGLE_PASS_LBVH), ("PLOC++", capi.PLOCPP), ("HPLOC", capi.HPLOC)):
                    try:
                        def best_of(graph):
                            for _ in range(3):
                                ctx.build(al, dm, n=mt.size, tris_on_device=True, use_graph=graph)
                            bst = None
                            for _ in range(10):
                                tt = ctx.build(al, dm, n=mt.size, tris_on_device=True, use_graph=graph)
                                tot = sum(tt.stage_ms[k] for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD))
                                if bst is None or tot < bst[0]:
                                    bst = (tot, [float(tt.stage_ms[k]) for k in (capi.T_EXTENTS, capi.T_MORTON, capi.T_SORT, capi.T_BUILD, capi.T_COLLAPSE)])
                            return bst, tt
                        plain, _ = best_of(False)
                        best, t = best_of(True)
                        res[nm] = {"total_ms": best[0], "Mprims_s": mt.size / best[0] / 1e3, "extents_morton_sort_build_collapse_ms": best[1],
                                   "launch": "CUDA graph replay", "total_ms_plain_launches": plain[0],
                                   "extents_morton_sort_build_collapse_ms_plain_launches": plain[1], "bvh4_cost": ctx.tree_cost(t)}
                        # primary rays on the built tree (TwoPassLbvh::traverseBvh, 512 x 512): the reference's four Bvh2 kernels and the Bvh4 walk
                        pr = TRACE_PRESETS[mesh]
                        tr = T.make_transform(pr["t"], pr["s"], [0.0, 0.0, 0.0, 1.0] if pr["q"] is None else qt_rotation(pr["q"]))
                        cam = T.make_camera(pr["eye"], qt_rotation(pr["cq"]), np.float32(45.0) * np.float32(np.pi) / np.float32(180.0))
                        d_rays, ray_ms = ctx.generate_rays(cam, 512, 512)
                        trace = {"ray_gen_ms": ray_ms}
                        for knm, kk in (("while_while", capi.TRAVERSE_WHILE), ("speculative_while", capi.TRAVERSE_SPECULATIVE_WHILE),
                                        ("if_if", capi.TRAVERSE_IFIF), ("restart_trail", capi.TRAVERSE_RESTART_TRAIL), ("bvh4", capi.TRAVERSE_WIDE4)):
                            tms = min(ctx.traverse(t, d_rays, 512 * 512, tr, kernel=kk)[2] for _ in range(5))
                            trace[knm] = {"ms": tms, "Mray_s": 512 * 512 / tms / 1e3}
                        ctx.free(d_rays)
                        res[nm]["primary_rays_512x512"] = trace
                        if nm in ("TwoPassLbvh", "PLOC++"):
                            # 512 x 512 rays are less than one full wave of a B200 (148 SMs x 2048 threads): the same view at 2048^2 and 4096^2
                            for side in (2048, 4096):
                                d_rays, _ = ctx.generate_rays(cam, side, side)
                                big = {}
                                for knm, kk in (("while_while", capi.TRAVERSE_WHILE), ("speculative_while", capi.TRAVERSE_SPECULATIVE_WHILE),
                                                ("if_if", capi.TRAVERSE_IFIF), ("restart_trail", capi.TRAVERSE_RESTART_TRAIL), ("bvh4", capi.TRAVERSE_WIDE4)):
                                    tms = min(ctx.traverse(t, d_rays, side * side, tr, kernel=kk)[2] for _ in range(3))
                                    big[knm] = {"ms": tms, "Mray_s": side * side / tms / 1e3}
                                ctx.free(d_rays)
                                res[nm][f"primary_rays_{side}x{side}"] = big
                    except capi.B2bvhError as e:
                        res[nm] = {"error": str(e)[:80]}
                ctx.free(dm)
                extras[mesh] = res
            line["reference_scenes"] = extras
            # ---- the sort stage against the same-box comparators: the reference's own (Orochi) kernels compiled unmodified, and CUB ----
            sc = os.path.join(ROOT, "baseline", "_ref", "sort_compare")
            if os.path.exists(sc):
                try:
                    r = subprocess.run([sc, capi.LIB_PATH, os.path.join(ROOT, "baseline", "_ref", "oro_radixsort.cubin"), str(n), str(TOTAL_100M)],
                                       capture_output=True, text=True, timeout=240)
                    line["sort_comparators"] = json.loads(r.stdout) if r.returncode == 0 else {"error": (r.stderr or r.stdout)[-200:]}
                except Exception as e:  # a bench tool: its failure must not cost the bench line
                    line["sort_comparators"] = {"error": str(e)[:200]}
            bunny = cpu_bunny(orc)
            if bunny:
                line["cpu_baseline_bunny"] = bunny

    if rank == 0:
        print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
