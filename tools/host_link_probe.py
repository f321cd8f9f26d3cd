"""Where does the end-to-end build time go on N GPUs?    bench.py's e2e step moves 640 MB up and 1.3 GB down per rank
over PCIe; this probe measures, with the same sizes and pinned buffers, the host<->device link of every rank ALONE (the ranks take
turns) and of ALL ranks at once — upload only, download only, both directions — and records the topology the box reports.  If the
all-at-once figures per rank fall well below the alone figures, the limit of e2e scaling is the host side (root complexes / memory),
not the builder.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_link_probe.py"""
import json
import os
import subprocess

import torch
import torch.distributed as dist

UP, DOWN = 640_000_000, 1_316_000_000


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    h_up = torch.empty(UP, dtype=torch.uint8).pin_memory()
    h_dn = torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    h_up.fill_(1)
    d_up = torch.empty(UP, dtype=torch.uint8, device="cuda")
    d_dn = torch.empty(DOWN, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once(up, dn):
        torch.cuda.synchronize()
        e0, e1, j = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
        e0.record(s1)
        s2.wait_event(e0)
        if up:
            with torch.cuda.stream(s1):
                d_up.copy_(h_up, non_blocking=True)
        if dn:
            with torch.cuda.stream(s2):
                h_dn.copy_(d_dn, non_blocking=True)
        j.record(s2)
        s1.wait_event(j)
        e1.record(s1)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def gbs(up, dn, reps=4):
        once(up, dn)
        ms = min(once(up, dn) for _ in range(reps))
        return ((UP if up else 0) + (DOWN if dn else 0)) / (ms * 1e-3) / 1e9

    def barrier():
        if world > 1:
            dist.barrier()

    res = {}
    for mode, (up, dn) in (("h2d", (True, False)), ("d2h", (False, True)), ("duplex", (True, True))):
        alone = 0.0
        for r in range(world):  # one rank at a time
            barrier()
            if r == rank:
                alone = gbs(up, dn)
        barrier()
        together = gbs(up, dn)  # every rank at once
        barrier()
        res[mode] = (alone, together)
    t = torch.tensor([v for k in ("h2d", "d2h", "duplex") for v in res[k]], device="cuda", dtype=torch.float64)
    allt = torch.zeros(world * t.numel(), device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_gather_into_tensor(allt, t)
    else:
        allt = t
    if rank == 0:
        rows = allt.reshape(world, 6).cpu().tolist()

        def sh(cmd):
            try:
                return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
            except Exception as e:
                return f"({e})"
        out = {"tool": "host_link_probe", "n_gpus": world, "bytes_up": UP, "bytes_down": DOWN,
               "per_rank_GBs": [{"rank": r, "h2d_alone": a, "h2d_all": b, "d2h_alone": c, "d2h_all": d, "duplex_alone": e, "duplex_all": f} for r, (a, b, c, d, e, f) in enumerate(rows)],
               "sum_all_ranks_GBs": {"h2d": sum(x[1] for x in rows), "d2h": sum(x[3] for x in rows), "duplex": sum(x[5] for x in rows)},
               "cpu_affinity_rank0": sorted(os.sched_getaffinity(0))[:4] + ["..."] + sorted(os.sched_getaffinity(0))[-1:], "cpus": os.cpu_count(),
               "nvidia_smi_topo": sh("nvidia-smi topo -m | head -20"), "numa": sh("numactl -H 2>/dev/null | head -12 || lscpu | grep -i numa"),
               "lscpu_numa": sh("lscpu | grep -i -E 'numa|model name|socket'")}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
