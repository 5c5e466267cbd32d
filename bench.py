#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched closed-chain state-validity hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload gd|pcs|rbs] [--samples S]

Default workload = BASELINE.json configs[1]: "GD projection batch: 1e6 random 12-DOF samples, KDL-Jacobian Newton closure
projection + FK + joint limits + collision, on 1 B200" (S-GD samples, env 1 / 3 poles).  One step = one pass of the hot
path (project -> limits -> link frames + collision) over one batch of S samples per GPU.  Weak scaling: every rank
(one process per GPU, torchrun) validates its own S samples per step; no data-path collective (samples are independent);
torch.distributed is used only for the barrier, the max-over-ranks of the device time and the summed counters.

Printed JSON line (rank 0): value = whole-job validated configurations / s with inputs resident in HBM, the steps issued on two caller
streams in turn (config.caller_streams; one_caller_stream = the same steps on one stream); the K steps are repeated
inside ONE timed region until it lasts >= 0.5 s (config.inner_repeats); e2e = the same metric through the host-buffer C-ABI
call (pinned host memory, H2D + D2H inside the timed region); roofline = the dominant kernel: k_gd_project against the
MEASURED FP64 FMA-chain peak (2600 flop x Newton iterations), k_collide (PCS / RBS) as an instruction-issue bound
FP32 kernel: executed box / triangle tests (its own counters) against the measured FP32 FMA-chain peak, issue-active from the
committed ncu capture; cpu_baseline = the reference's CPU path timed on this box (bounded sample, 1 core).
The default run also carries, under "extra", complete sub-records (value, e2e, roofline, cpu_baseline, clocks) for the two other
BASELINE workloads -- "pcs" (configs[4], PCS arm; CPU arm = the reference's own compiled code) and "rbs" (configs[2]: 1e5 edges,
checkMotionRBS) -- and "planning" (configs[3]): single-query latency and 500-query throughput of the restated planners.
`--impl reference` times the reference's CPU path on all host cores (GD: the restated KDL loop, kind "port" -- real KDL is
not installable; PCS / RBS: the reference's own compiled code through oracle/_ref, kind "reference").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

NB = 4     # distinct resident input batches cycled through the timed steps (aggregate footprint > L2)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="gd", choices=["gd", "pcs", "rbs"])
    ap.add_argument("--samples", type=int, default=0, help="units per GPU per step (default: 1e6 samples; 1e5 edges for rbs = BASELINE configs[2])")
    ap.add_argument("--streams", type=int, default=2, help="caller streams the resident GD steps alternate over (2: step i+1's projection overlaps step i's collision tail)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary PCS / RBS measurements")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------------------------
# clocks sampled during the timed region
# ------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------------
def make_rbs_edges(ck, sampling, n_edges, seed):
    """S-RBS protocol (tests/test_rbs_pcs.cpp:39-47): A = valid S-PCS sample, B = A + s (U - A), s in [0.01, 1], projected on
    A's (ac = 0, ik) and valid.  Built with the GPU path itself (host API)."""
    import numpy as np
    qa, qb, iks = [], [], []
    i0, have = 0, 0
    lo = np.array([-165, -110, -110, -160, -120]) * 3.1416 / 180
    hi = np.array([165, 110, 70, 160, 120]) * 3.1416 / 180
    lo = np.append(lo, -3.1416); hi = np.append(hi, 3.1416)
    while have < n_edges:
        m = 1 << 21
        q, ik = sampling.sample_pcs(seed, i0, m)
        ok, valid, qo = ck.pcs_validate(q, 0, ik)
        idx = np.nonzero(valid)[0]
        a = qo[idx]
        u = np.stack([sampling.uniform(seed + 1000, i0, m, k)[idx] for k in range(6)], axis=1)
        s = 0.01 + 0.99 * sampling.uniform(seed + 1000, i0, m, 6)[idx]
        b = a.copy()
        b[:, :6] = a[:, :6] + s[:, None] * ((lo + u * (hi - lo)) - a[:, :6])
        okb, vb, qbo = ck.pcs_validate(b, 0, ik[idx])
        keep = vb == 1
        qa.append(a[keep]); qb.append(qbo[keep]); iks.append(ik[idx][keep])
        have += int(keep.sum())
        i0 += m
    import numpy as np
    return np.concatenate(qa)[:n_edges], np.concatenate(qb)[:n_edges], np.concatenate(iks)[:n_edges]


def bind_to_gpu_numa_node(local_rank, world):
    """Pin this rank (and therefore the first touch of its pinned host buffers and its copy threads) to host cores of its own:
    the CPUs of the GPU's NUMA node when the box has several nodes, else an equal slice of the allowed cores per rank (on a
    single-node box all ranks share one memory system; distinct cores at least keep the ranks' host threads from migrating onto
    each other).  Returns a dict that goes into config (what was done and why)."""
    info = {"numa_node": None, "cpus": None, "how": None}
    allowed = sorted(os.sched_getaffinity(0))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node >= 0 and len(nodes) > 1:
            cpus = set()
            for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
            mine = sorted(set(allowed) & cpus)
            if mine:
                os.sched_setaffinity(0, mine)
                info.update(numa_node=node, cpus="%d-%d (%d)" % (mine[0], mine[-1], len(mine)), how="cpus of the GPU's NUMA node")
                return info
        info["numa_node"] = node
        info["how"] = "single NUMA node (node %d of %d)" % (node, len(nodes))
    except Exception as e:                      # no pynvml / sysfs entry: say so instead of hiding it
        info["how"] = "numa lookup failed: %r" % (e,)
    if world > 1 and len(allowed) >= 2 * world:
        per = len(allowed) // world
        mine = allowed[local_rank * per:(local_rank + 1) * per]
        try:
            os.sched_setaffinity(0, mine)
            info["cpus"] = "%d-%d (%d)" % (mine[0], mine[-1], len(mine))
            info["how"] = (info["how"] or "") + "; equal slice of the allowed cores per rank"
        except Exception as e:
            info["how"] = (info["how"] or "") + "; sched_setaffinity failed: %r" % (e,)
    return info


# flop of one conservative box test (csrc/ckc_device.cuh boxes_farther_than: two box_to_world 2 x 57, centre difference 3, t 15, R 45,
# |R| 9, three + three face axes 24 + 39, nine edge-cross axes 135) and of one exact triangle-pair distance
FLOP_BOX_TEST = 384.0
FLOP_TRI_TEST = 1000.0
MIN_TIMED_S = 0.5          # the timed region repeats the K steps until it is at least this long


def collide_roofline(stats, kms, kl, peak32, peak64, stage_share, extra=None, capture="k_collide"):
    """k_collide: instruction-issue / latency bound FP32 + integer code (ncu: profiles/r2_ncu_collide.txt), not an FP64 kernel.  achieved =
    the work the kernel EXECUTED by its own counters (box tests x 384 flop + exact triangle tests x 1000 flop) / its CUDA-event time,
    against the measured FP32 FMA-chain peak.  (The reference-work model of SURVEY 8d -- PQP's 116 full queries per configuration --
    is not used for a fraction: the kernel prunes most of that work away.)"""
    dur = kms / max(kl, 1) * 1e-3
    flop = (FLOP_BOX_TEST * stats["bv_tests"] + FLOP_TRI_TEST * stats["tri_tests"]) / max(kl, 1)
    ach = flop / dur / 1e12 if dur > 0 else 0.0
    ncu = {}
    try:
        ncu = json.load(open(os.path.join(REPO, "profiles", "r2_traffic.json")))[capture]
    except Exception:
        pass
    r = {"bound": "issue/fp32", "kernel": "k_collide", "achieved": ach, "peak": peak32, "unit": "TFLOP/s", "frac": ach / peak32 if peak32 else None,
         "traffic": ncu.get("traffic"), "traffic_source": ncu.get("note"),
         "issue_active_pct_ncu": ncu.get("issue_active_pct"), "lanes_active_ncu": ncu.get("lanes_per_instruction"),
         "peak_source": "FP32 FMA-chain peak measured live by ckc_measure_fma_peak (FP64: %.1f TFLOP/s); MEASURED_PEAKS.json carries HBM / bf16 only" % peak64,
         "work_model": "executed work by the kernel's own counters: %.0f flop per box test, %.0f per exact triangle test" % (FLOP_BOX_TEST, FLOP_TRI_TEST),
         "kernel_ms_per_launch": kms / max(kl, 1), "collision_checks_per_launch": stats["collision_calls"] / max(kl, 1),
         "stage_share_of_step": stage_share}
    if extra:
        r.update(extra)
    return r


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from abb_dualarm_planning_b200 import Checker, sampling

    rank, local_rank, world = dist_env()
    binding = bind_to_gpu_numa_node(local_rank, world)
    if args.gpus > 1 and world != args.gpus:
        raise SystemExit("launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ck = Checker(env=1, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    K, W = args.steps, max(args.warmup, 0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from abb_dualarm_planning_b200.sharding import reduce_scalar

    def max_over_ranks(x):           # the only cross-rank traffic of a run: timing agreement and summed counters, never data
        return reduce_scalar(x, "max", dist, dev)

    def sum_over_ranks(x):
        return reduce_scalar(x, "sum", dist, dev)

    def repeats_for(step_fn, timer):
        """inner repeats R of the K steps so that the timed region lasts >= MIN_TIMED_S (same R on every rank)"""
        t = timer(lambda: [step_fn(k % NB) for k in range(min(K, 4))]) / min(K, 4)          # seconds per step, untimed estimate
        t = max_over_ranks(t)
        return max(1, int(np.ceil(MIN_TIMED_S / max(K * t, 1e-9))))

    side = []                        # extra caller streams of the resident GD steps (--streams), empty otherwise

    def fork_streams():              # the side streams start after everything issued so far on the current stream ...
        if side:
            ev = torch.cuda.Event(); ev.record()
            for st_ in side: st_.wait_event(ev)

    def join_streams():              # ... and the current stream continues after everything issued on them
        for st_ in side:
            ev = torch.cuda.Event(); ev.record(st_); torch.cuda.current_stream().wait_event(ev)

    def ev_timer(fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fork_streams(); fn(); join_streams(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3

    def wall_timer(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        return time.perf_counter() - t0

    peaks = {}

    def peak(p):
        if p not in peaks:
            peaks[p] = ck.fma_peak(p)
        return peaks[p]

    def measure_projection(wl, n):
        """S-GD / S-PCS: value (inputs resident in HBM), e2e (host-buffer C-ABI call), roofline of the dominant kernel"""
        out = {}
        seeds = [1000 * (rank + 1) + b for b in range(NB)]
        host_q, host_ik, d_q, d_ik = [], [], [], []
        for s in seeds:
            if wl == "gd":
                q = sampling.sample_gd(s, 0, n); ik = None
            else:
                q, ik = sampling.sample_pcs(s, 0, n)
            hq = torch.from_numpy(q).pin_memory()
            host_q.append(hq)
            d_q.append(hq.to(dev, non_blocking=True).t().contiguous())          # SoA [12][n]
            if ik is not None:
                hik = torch.from_numpy(ik).pin_memory(); host_ik.append(hik); d_ik.append(hik.to(dev))
        d_qout = torch.empty((12, n), dtype=torch.float64, device=dev)
        d_ok = torch.empty(n, dtype=torch.uint8, device=dev)
        d_valid = torch.empty(n, dtype=torch.uint8, device=dev)
        # caller streams of the resident GD steps: with 2, consecutive steps are issued on alternating streams with their own output buffers
        NS = max(1, args.streams)
        side[:] = [torch.cuda.Stream(device=dev) for _ in range(NS)] if NS > 1 else []
        outs = [(d_qout, d_valid)] + [(torch.empty_like(d_qout), torch.empty_like(d_valid)) for _ in range(NS - 1)]
        torch.cuda.synchronize()
        step_no = [0]

        def step_dev(b):
            if wl == "gd":
                slot = step_no[0] % len(side) if side else 0
                step_no[0] += 1
                qo, va = outs[slot]
                ck.gd_validate_dev(n, d_q[b].data_ptr(), qo.data_ptr(), va.data_ptr(), side[slot].cuda_stream if side else stream)
            else:
                slot = step_no[0] % len(side) if side else 0
                step_no[0] += 1
                qo, va = outs[slot]
                ck.pcs_validate_dev(n, d_q[b].data_ptr(), None, d_ik[b].data_ptr(), qo.data_ptr(), d_ok.data_ptr(), va.data_ptr(), side[slot].cuda_stream if side else stream)

        for w in range(W):
            step_dev(w % NB)
        barrier()
        R = repeats_for(step_dev, ev_timer)
        barrier()
        ck.reset_stats()
        sampler = ClockSampler(local_rank); sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        fork_streams()
        for r in range(R):
            for k in range(K):
                step_dev(k % NB)
        join_streams()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        out["clocks"] = sampler.stop()
        stats = ck.stats()
        steps_run = K * R
        if side:                         # for the record: the same K steps issued on ONE caller stream (every call waits for the one before it)
            keep = list(side); side[:] = []
            barrier()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            for r in range(max(1, R // 2)):
                for k in range(K):
                    step_dev(k % NB)
            e3.record()
            barrier()
            ms1 = max_over_ranks(e2.elapsed_time(e3))
            out["one_caller_stream"] = {"value": world * K * max(1, R // 2) * n / (ms1 * 1e-3), "ms_per_step": ms1 / (K * max(1, R // 2))}
            out["caller_streams"] = len(keep)
            side[:] = keep
        # per-kernel durations for the roofline: K steps once more with the chunk overlap switched off, so that every kernel runs
        # alone on the stream and its CUDA-event bracket measures that kernel only
        torch.cuda.synchronize()
        side[:] = []                     # the serial pass and everything after it run on the current stream
        ck.set_overlap(False); ck.set_stage_timing(True)
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record()
        for k in range(K):
            step_dev(k % NB)
        es1.record()
        torch.cuda.synchronize()
        ms_serial = es0.elapsed_time(es1)
        stages = ck.stage_times()
        ck.set_stage_timing(False); ck.set_overlap(True)
        n_valid_last = int(d_valid.sum().item())
        out["value"] = world * steps_run * n / (ms * 1e-3)
        out["ms_per_step"] = ms / steps_run
        out["timed_region_s"] = ms * 1e-3
        out["inner_repeats"] = R
        out["gpu_launches"] = int(sum_over_ranks(stats["kernel_launches"]))

        # ---- e2e: host-buffer C-ABI call, pinned host memory, H2D + compute + D2H inside the timed region ----
        import ctypes as C
        h_qout = torch.empty((n, 12), dtype=torch.float64).pin_memory()
        h_valid = torch.empty(n, dtype=torch.uint8).pin_memory()
        h_ok = torch.empty(n, dtype=torch.uint8).pin_memory()
        cap = n // 8
        h_idx = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_qv = torch.empty((cap, 12), dtype=torch.float64).pin_memory()
        nv = C.c_int64()

        def step_host_full(b):
            if wl == "gd":
                ck._chk(ck.L.ckc_gd_validate(ck.ctx, n, host_q[b].data_ptr(), h_qout.data_ptr(), h_valid.data_ptr()))
            else:
                ck._chk(ck.L.ckc_pcs_validate(ck.ctx, n, host_q[b].data_ptr(), None, host_ik[b].data_ptr(), h_qout.data_ptr(), h_ok.data_ptr(), h_valid.data_ptr()))

        def step_host_compact(b):
            if wl == "gd":
                ck._chk(ck.L.ckc_gd_validate_compact(ck.ctx, n, host_q[b].data_ptr(), h_valid.data_ptr(), C.byref(nv), h_idx.data_ptr(), h_qv.data_ptr(), cap))
            else:
                ck._chk(ck.L.ckc_pcs_validate_compact(ck.ctx, n, host_q[b].data_ptr(), None, host_ik[b].data_ptr(), h_valid.data_ptr(), C.byref(nv),
                                                      h_idx.data_ptr(), h_qv.data_ptr(), cap))
        e2e = {}
        for name, fn in (("full", step_host_full), ("compact", step_host_compact)):
            for w in range(min(W, 2) or 1):
                fn(w % NB)
            barrier()
            Re = repeats_for(fn, wall_timer)
            barrier()
            t0 = time.perf_counter()
            for r in range(Re):
                for k in range(K):
                    fn(k % NB)
            torch.cuda.synchronize()
            e2e[name] = max_over_ranks(time.perf_counter() - t0) / Re
        h2d = n * 96 + (n if wl == "pcs" else 0)
        out["e2e"] = {"value": world * K * n / e2e["compact"], "unit": "configs/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": n + 8 + d2h_compact_bytes(n),
                      "api": "ckc_%s_validate_compact (mask + index and projected configuration of the valid samples), pinned host buffers" % wl,
                      "valid_last_step": int(nv.value), "checksum_valid_last_step": int(h_valid.sum().item()),
                      "full_output": {"value": world * K * n / e2e["full"], "d2h_bytes_per_step": n * 96 + n + (n if wl == "pcs" else 0),
                                      "api": "ckc_%s_validate (every projected configuration returned)" % wl}}

        # ---- roofline of the dominant kernel (rank 0's own numbers) ----
        per_step = steps_run * n
        share = {k: round(v["ms"] / max(ms_serial, 1e-9), 4) for k, v in stages.items()}
        timing_note = ("kernel durations from a second pass of the same K steps with ckc_set_overlap(0) (every kernel alone on the stream, %.3f ms/step); "
                       "the timed region overlaps chunks (%.3f ms/step)" % (ms_serial / K, ms / steps_run))
        serial_stats = ck.stats()            # counters since the reset: timed region + the serial pass (K more steps)
        if wl == "gd":
            iters = stats["gd_iterations"] / per_step
            flop_launch = iters * 2600.0 * n                               # 2600 flop per Newton iteration
            kms, kl = stages["project"]["ms"], stages["project"]["launches"]
            dur = kms / max(kl, 1) * 1e-3
            achieved = flop_launch / dur / 1e12 if dur > 0 else 0.0
            traffic, traffic_src, ncu = None, None, {}
            try:
                ncu = json.load(open(os.path.join(REPO, "profiles", "r2_traffic.json")))["k_gd_project"]
                traffic, traffic_src = ncu["traffic"], ncu["note"]
            except Exception:
                pass
            p64 = peak(64)
            out["roofline"] = {"bound": "fp64", "kernel": "k_gd_project", "achieved": achieved, "peak": p64, "unit": "TFLOP/s",
                               "frac": achieved / p64 if p64 else None, "traffic": traffic, "traffic_source": traffic_src,
                               "fp64_pipe_active_pct_ncu": ncu.get("fp64_pipe_active_pct"),
                               "peak_source": "FP64 FMA-chain peak measured live by ckc_measure_fma_peak (MEASURED_PEAKS.json has HBM/bf16 only; FP32 FMA peak %.1f TFLOP/s)" % peak(32),
                               "work_model": "algorithmic flop of the reference path (SURVEY 8d): 2600 flop x measured Newton iterations (%.3f per sample)" % iters,
                               "kernel_ms_per_launch": kms / max(kl, 1), "stage_share_of_step": share, "timing_note": timing_note,
                               "hbm_gbs_algorithmic": (per_step * 200.0) / (ms * 1e-3) / 1e9, "hbm_peak_gbs_measured": measured_hbm_peak()}
        else:
            # the serial pass alone: counters of the whole region scaled to the K serial steps
            sc = {k: serial_stats[k] * K / (steps_run + K) for k in ("bv_tests", "tri_tests", "collision_calls")}
            out["roofline"] = collide_roofline(sc, stages["collide"]["ms"], stages["collide"]["launches"], peak(32), peak(64), share,
                                               {"timing_note": timing_note, "hbm_gbs_algorithmic": (per_step * 200.0) / (ms * 1e-3) / 1e9,
                                                "hbm_peak_gbs_measured": measured_hbm_peak()}, capture="k_collide_pcs")
        out["path_stats"] = {"newton_iters_mean": stats["gd_iterations"] / per_step if wl == "gd" else None,
                             "fraction_passing_projection": stats["ik_feasible"] / per_step,
                             "collision_checks_per_step": stats["collision_calls"] / steps_run, "bv_tests_per_check": stats["bv_tests"] / max(stats["collision_calls"], 1),
                             "tri_tests_per_check": stats["tri_tests"] / max(stats["collision_calls"], 1), "valid_last_step": n_valid_last,
                             "configs_sent_to_heavy_kernel": stats["queue_overflows"]}
        del d_q, d_ik, d_qout, host_q, host_ik
        return out

    def measure_rbs(n):
        out = {}
        qa, qb, iks = make_rbs_edges(ck, sampling, n, 7000 + rank)
        # pinned host buffers for the edge set and the results (the contract's e2e: inputs from pinned host memory every step)
        h_qa = torch.from_numpy(np.ascontiguousarray(qa)).pin_memory(); h_qb = torch.from_numpy(np.ascontiguousarray(qb)).pin_memory()
        h_ik = torch.from_numpy(np.ascontiguousarray(iks)).pin_memory()
        h_ok = torch.empty(n, dtype=torch.uint8).pin_memory(); h_nc = torch.empty(n, dtype=torch.int32).pin_memory()

        def rbs_call():
            ck._chk(ck.L.ckc_pcs_check_motion_rbs(ck.ctx, n, h_qa.data_ptr(), h_qb.data_ptr(), None, h_ik.data_ptr(), h_ok.data_ptr(), h_nc.data_ptr()))
        for w in range(max(W, 1)):
            rbs_call()
        barrier()
        t1 = max_over_ranks(wall_timer(rbs_call))
        R = max(1, int(np.ceil(MIN_TIMED_S / max(K * t1, 1e-9))))
        ck.reset_stats(); ck.set_stage_timing(True)
        sampler = ClockSampler(local_rank); sampler.start()
        barrier()
        t0 = time.perf_counter()
        okc = 0
        for r in range(R):
            for k in range(K):
                rbs_call()
        okc = int(h_ok.sum().item())
        torch.cuda.synchronize()
        t = max_over_ranks(time.perf_counter() - t0)
        out["clocks"] = sampler.stop()
        stats = ck.stats(); stages = ck.stage_times(); ck.set_stage_timing(False)
        steps_run = K * R
        out["value"] = world * steps_run * n / t
        out["ms_per_step"] = 1e3 * t / steps_run
        out["timed_region_s"] = t
        out["inner_repeats"] = R
        out["gpu_launches"] = int(sum_over_ranks(stats["kernel_launches"]))
        out["e2e"] = {"value": out["value"], "unit": "edges/s", "h2d_bytes_per_step": n * 194, "d2h_bytes_per_step": n * 5,
                      "note": "the RBS entry point only exists in host-buffer form (pinned host buffers in, flags + check counts out; the frontier lives on the device); value == e2e"}
        share = {k: round(v["ms"] / (1e3 * t), 4) for k, v in stages.items()}
        out["roofline"] = collide_roofline(stats, stages["collide"]["ms"], stages["collide"]["launches"], peak(32), peak(64), share,
                                           {"timing_note": "one k_collide launch per bisection level (%.1f levels per call); host loop + level synchronisation = %.0f %% of the wall time"
                                            % (stages["collide"]["launches"] / steps_run, 100.0 * (1 - sum(v["ms"] for v in stages.values()) / (1e3 * t)))})
        out["path_stats"] = {"edges_ok_fraction": okc / n, "checks_per_edge": stats["collision_calls"] / (steps_run * n),
                             "levels_per_call": stages["collide"]["launches"] / steps_run}
        return out

    METRIC = {"gd": "validated closed-chain configs/s (project + FK/limits + collide)", "pcs": "validated closed-chain configs/s (project + FK/limits + collide)",
              "rbs": "RBS edges/s (checkMotionRBS, PCS)"}

    def workload_text(wl, n):
        return {"gd": "BASELINE configs[1]: S-GD, %d uniform 12-DOF samples per GPU per step, KDL-style Newton closure projection + joint limits + 116-pair mesh collision, env 1 (3 poles)" % n,
                "pcs": "BASELINE configs[4] (PCS arm): S-PCS, %d samples per GPU per step, APC/PCS analytic IK projection + 116-pair mesh collision, env 1 (3 poles)" % n,
                "rbs": "BASELINE configs[2]: S-RBS, %d edges per GPU per step, checkMotionRBS (PCS), RBS_tol 0.05, env 1 (3 poles)" % n}[wl]

    wl = args.workload
    n = args.samples or (100000 if wl == "rbs" else 1_000_000)
    out = measure_rbs(n) if wl == "rbs" else measure_projection(wl, n)

    # ---- CPU baseline (rank 0, N = 1 only) ----
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(wl, cores=1, budget_s=12.0)

    # ---- the other two workloads as complete sub-records (same run, same box): BASELINE configs[2] and the PCS arm of configs[4] ----
    extra = {}
    if not args.no_extra and wl == "gd":
        for owl in ("pcs", "rbs"):
            on = 100000 if owl == "rbs" else 1_000_000
            rec = measure_rbs(on) if owl == "rbs" else measure_projection(owl, on)
            rec.update(metric=METRIC[owl], unit="edges/s" if owl == "rbs" else "configs/s", n_gpus=world, steps=K, warmup=W, higher_is_better=True, scaling="weak",
                       dtype="f64", data="synthetic", config={"workload": workload_text(owl, on), "units_per_gpu_per_step": on, "inner_repeats": rec["inner_repeats"]})
            if rank == 0 and world == 1 and not args.no_cpu:
                rec["cpu_baseline"] = cpu_baseline(owl, cores=1, budget_s=6.0)
            extra[owl] = rec
        if rank == 0 and world == 1:
            extra.update(extra_measurements(ck, sampling, torch, dev, stream))

    if rank == 0:
        line = {"metric": METRIC[wl], "value": out["value"], "unit": "configs/s" if wl != "rbs" else "edges/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": out["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text(wl, n), "units_per_gpu_per_step": n, "parallelism": "shard%d (independent samples, no collective)" % world,
                           "host_binding_rank0": binding, "inner_repeats": out["inner_repeats"], "timed_region_s": out["timed_region_s"],
                           "timing": "the K steps are repeated inner_repeats times inside ONE timed region (>= %.1f s); ms_per_step = region / (K x inner_repeats)" % MIN_TIMED_S,
                           "cache": "%d distinct resident input batches cycled (%.0f MB aggregate > 126 MB L2)" % (NB, NB * n * 96 / 1e6) if wl != "rbs" else "edge set re-uploaded from host every step"},
                "e2e": out["e2e"], "gpu_launches": out["gpu_launches"], "clocks": out["clocks"], "roofline": out["roofline"],
                "path_stats": out.get("path_stats")}
        if "caller_streams" in out:
            line["config"]["caller_streams"] = out["caller_streams"]
            line["config"]["caller_streams_note"] = ("the resident steps are issued on %d caller streams in turn (each with its own output buffers; the library rotates its scratch "
                                                     "lanes per call): step i+1's projection kernel fills the tail of step i's collision kernel; one_caller_stream = the same "
                                                     "steps on one stream" % out["caller_streams"])
            line["one_caller_stream"] = out["one_caller_stream"]
        if "cpu_baseline" in out:
            line["cpu_baseline"] = out["cpu_baseline"]
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    ck.close()
    if world > 1:
        dist.destroy_process_group()


def _planner_run(exe, env, nq, step, flav, workers, td):
    outp = os.path.join(td, "paths_%s_%d.txt" % (flav, nq))
    subprocess.run([exe, "1", str(nq), step, "7", outp, str(workers), flav], env=env, capture_output=True, timeout=600, check=True)
    t = open(outp).read().strip().split("\n")[-1].split()
    return dict(zip(t[1::2], t[2::2]))


def extra_measurements(ck, sampling, torch, dev, stream):
    """S-PCS generated on the device, and BASELINE configs[3] (planning query): latency of ONE query and throughput of 500"""
    ex = {}
    n = 1 << 24
    d_ok = torch.empty(n, dtype=torch.uint8, device=dev); d_valid = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(2):
        ck.spcs_validate_seeded_dev(99, 0, n, None, d_ok.data_ptr(), d_valid.data_ptr(), stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 8
    two = [torch.cuda.Stream(device=dev) for _ in range(2)]                      # two caller streams with their own mask buffers, as in the resident steps
    d_ok2 = torch.empty_like(d_ok); d_valid2 = torch.empty_like(d_valid)
    e0.record()
    for st_ in two: st_.wait_event(e0)
    for r in range(reps):
        ck.spcs_validate_seeded_dev(100 + r, 0, n, None, (d_ok if r % 2 == 0 else d_ok2).data_ptr(), (d_valid if r % 2 == 0 else d_valid2).data_ptr(), two[r % 2].cuda_stream)
    for st_ in two:
        ev = torch.cuda.Event(); ev.record(st_); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    ex["pcs_seeded_configs_per_s"] = reps * n / (e0.elapsed_time(e1) * 1e-3)
    ex["pcs_seeded_note"] = "S-PCS samples generated on the device from the counter-based generator (no input traffic), %d samples x %d" % (n, reps)
    del d_ok, d_valid, d_ok2, d_valid2
    try:
        import tempfile
        exe = os.path.join(REPO, "abb_dualarm_planning_b200", "host", "cbirrt_pcs_batch")
        if os.path.exists(exe):
            with tempfile.TemporaryDirectory() as td:
                env = dict(os.environ, CKC_MESH_PATH=os.path.join(REPO, "abb_dualarm_planning_b200", "data", "meshes.ckcmesh"),
                           CKC_DEVICE=str(torch.cuda.current_device()))
                plan = {}
                for flav, step, ref_ms in (("pcs", "1.0", 218), ("gd", "2.6", 128), ("hb", "1.4", 342)):
                    one = _planner_run(exe, env, 40, step, flav, 0, td)           # workers = 0: 40 queries one after another, each alone on the device
                    many = _planner_run(exe, env, 500, step, flav, 4, td)
                    plan["cbirrt_" + flav] = {
                        "latency_ms_1_query": float(one["latency_mean_ms"]), "latency_median_ms_1_query": float(one["latency_median_ms"]),
                        "latency_max_ms_1_query": float(one["latency_max_ms"]), "queries_for_latency": int(one["queries"]), "solved_for_latency": int(one["solved"]),
                        "gpu_calls_per_query": int(one["gpu_calls"]) / max(int(one["queries"]), 1), "planner_steps_per_query": float(one["mean_rounds"]),
                        "throughput_ms_per_query_500": float(many["ms_per_query"]), "solved_of_500": int(many["solved"]), "seconds_500": float(many["seconds"]),
                        "reference_published_ms_per_query": ref_ms,
                        "note": "latency = mean wall time of one env-1 start-to-goal query run ALONE on the device (40 seeds, one after another) -- what BASELINE.md's %d ms "
                                "measures (mean of 500 runs, there on the author's CPU); throughput = wall time of 500 concurrent queries / 500 (lock-step ensemble, "
                                "host/cbirrt_pcs_batch.cpp, 4 workers) -- a different quantity, not comparable with the published latency" % ref_ms}
                    if flav == "pcs":          # batching inside ONE query: a growTree call's run of steps as one window (same tree as step by step)
                        sbs = _planner_run(exe, dict(env, CKC_PLAN_WINDOW="1"), 40, step, flav, 0, td)
                        plan["cbirrt_pcs"].update({"window_steps_per_round": 64, "latency_ms_1_query_step_by_step": float(sbs["latency_mean_ms"]),
                                                   "gpu_calls_per_query_step_by_step": int(sbs["gpu_calls"]) / max(int(sbs["queries"]), 1)})
                lexe = os.path.join(REPO, "abb_dualarm_planning_b200", "host", "lazy_gd_batch")
                if os.path.exists(lexe):           # BASELINE configs[3] names SBL_GD: planners/SBL_GD.cpp restated over the C ABI, max step 0.8 (the best step of BASELINE.md)
                    def lazy(nq, workers):
                        outp = os.path.join(td, "sbl_%d.txt" % nq)
                        subprocess.run([lexe, "sbl", "1", str(nq), "0.8", "7", outp, str(workers)], env=env, capture_output=True, timeout=600, check=True)
                        t = open(outp).read().strip().split("\n")[-1].split()
                        return dict(zip(t[1::2], t[2::2]))
                    one, many = lazy(40, 0), lazy(500, 4)
                    plan["sbl_gd"] = {"latency_ms_1_query": float(one["latency_mean_ms"]), "latency_median_ms_1_query": float(one["latency_median_ms"]),
                                      "latency_max_ms_1_query": float(one["latency_max_ms"]), "queries_for_latency": int(one["queries"]), "solved_for_latency": int(one["solved"]),
                                      "gpu_calls_per_query": int(one["gpu_calls"]) / max(int(one["queries"]), 1),
                                      "throughput_ms_per_query_500": float(many["ms_per_query"]), "solved_of_500": int(many["solved"]),
                                      "reference_published_ms_per_query": 163,
                                      "note": "SBL_GD restated (host/lazy_gd_batch.cpp), max step 0.8 (the reference's best, BASELINE.md: 163 ms per query on the author's CPU); latency = one query alone, mean of 40 seeds"}
                ex["planning"] = plan
    except Exception as e:       # an extra, never the metric
        ex["planning"] = {"not_run": repr(e)}
    return ex


def measured_hbm_peak():
    """HBM copy bandwidth of this pool's B200 from the driver-written MEASURED_PEAKS.json (GB/s); the profiling recipe's figure if absent"""
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6545.9


def d2h_compact_bytes(n):
    """bytes the compact entry points copy back besides the mask: per pipeline chunk a count and a fixed-size prefix (5 % of the
    chunk + 1024 entries of index + configuration), the chunking of project_host in csrc/ckc_host.cu"""
    c = max(1, min(1 << 18, max(min(n, 1 << 16), (n + 3) // 4)))
    nchunks = (n + c - 1) // c
    return nchunks * (4 + (c // 20 + 1024) * 100)


# ------------------------------------------------------------------------------------------------------------------
# CPU side: baseline / reference arm
# ------------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """one process: validates samples [i0, i0+m) of the workload with the CPU restatement; returns (seconds, n_valid)"""
    wl, seed, i0, m = args
    import numpy as np
    from abb_dualarm_planning_b200 import sampling
    from oracle import lib as olib
    orc = olib.Oracle(1)
    if wl == "gd":
        q = sampling.sample_gd(seed, i0, m)
        t0 = time.perf_counter()
        valid, _ = orc.gd_validate(q)
        return time.perf_counter() - t0, int(valid.sum())
    if wl == "pcs":
        q, ik = sampling.sample_pcs(seed, i0, m)
        t0 = time.perf_counter()
        ok, valid, _ = orc.pcs_validate(q, 0, ik)
        return time.perf_counter() - t0, int(valid.sum())
    raise ValueError(wl)


def _ref_worker(args):
    """one process: the reference's OWN compiled code (oracle/_ref: prebuilt tests/trbspcs through the LD_PRELOAD driver)"""
    wl, seed, i0, m = args
    from oracle import refshim
    ok, valid, secs = refshim.spcs_seeded(seed, i0, m)
    return secs, int(valid.sum())


def _cpu_rbs_edges(orc, seed, n_edges):
    """S-RBS edges built with the CPU oracle (untimed): valid endpoint A, B = A + s (U - A) projected on A's branch and valid"""
    import numpy as np
    from abb_dualarm_planning_b200 import sampling
    lo = np.append(np.array([-165, -110, -110, -160, -120]) * 3.1416 / 180, -3.1416)
    hi = np.append(np.array([165, 110, 70, 160, 120]) * 3.1416 / 180, 3.1416)
    qa, qb, iks, i0, have = [], [], [], 0, 0
    while have < n_edges:
        m = 200000
        q, ik = sampling.sample_pcs(seed, i0, m)
        ok, valid, qo = orc.pcs_validate(q, 0, ik)
        idx = np.nonzero(valid)[0]
        a = qo[idx]
        u = np.stack([sampling.uniform(seed + 1000, i0, m, k)[idx] for k in range(6)], axis=1)
        s = 0.01 + 0.99 * sampling.uniform(seed + 1000, i0, m, 6)[idx]
        b = a.copy(); b[:, :6] = a[:, :6] + s[:, None] * ((lo + u * (hi - lo)) - a[:, :6])
        okb, vb, qbo = orc.pcs_validate(b, 0, ik[idx])
        keep = vb == 1
        qa.append(a[keep]); qb.append(qbo[keep]); iks.append(ik[idx][keep]); have += int(keep.sum()); i0 += m
    return np.concatenate(qa)[:n_edges], np.concatenate(qb)[:n_edges], np.concatenate(iks)[:n_edges]


def _rbs_worker(args):
    """one process: checkMotionRBS on its own S-RBS edges -- the reference's compiled code when oracle/_ref exists, else the port"""
    seed, n_edges, use_ref = args
    from oracle import lib as olib, refshim
    orc = olib.Oracle(1)
    qa, qb, iks = _cpu_rbs_edges(orc, seed, n_edges)
    if use_ref:
        ok, nodes, secs = refshim.rbs(qa, qb, 0, iks, count_nodes=False)
        return secs, int(ok.sum())
    t0 = time.perf_counter()
    ok, nc = orc.pcs_check_motion_rbs(qa, qb, 0, iks)
    return time.perf_counter() - t0, int(ok.sum())


def cpu_baseline(wl, cores, budget_s):
    """bounded sample of the same workload on `cores` host processes"""
    import multiprocessing as mp
    from oracle import refshim
    if wl == "rbs":
        use_ref = refshim.available()
        m = max(int(budget_s / (3.0e-3 if use_ref else 0.4e-3)), 50)
        jobs = [(7000 + c, m, use_ref) for c in range(cores)]
        t0 = time.perf_counter()
        if cores == 1:
            res = [_rbs_worker(jobs[0])]
        else:
            with mp.get_context("fork").Pool(cores) as pool:
                res = pool.map(_rbs_worker, jobs)
        wall = time.perf_counter() - t0
        busy = max(r[0] for r in res)
        return {"value": cores * m / busy, "unit": "edges/s", "cores": cores, "kind": "reference" if use_ref else "port",
                "sample": "%d S-RBS edges per core, checkMotionRBS compute time of the slowest core %.2f s (edge construction untimed, wall %.2f s); %s" % (
                    m, busy, wall, "the reference's own compiled code (tests/trbspcs via oracle/_ref)" if use_ref else "oracle/ C++ restatement"),
                "n_ok": sum(r[1] for r in res)}
    wl_eff = wl
    use_ref = wl_eff == "pcs" and refshim.available()
    # per-unit CPU cost (measured on this image): GD port ~90 us, PCS port ~3 us, PCS reference binary ~36 us
    per_unit = {"gd": 95e-6, "pcs": 36e-6 if use_ref else 3.5e-6}[wl_eff]
    m = max(int(budget_s / per_unit), 1000)
    jobs = [(wl_eff, 1001, c * m, m) for c in range(cores)]
    fn = _ref_worker if use_ref else _cpu_worker
    t0 = time.perf_counter()
    if cores == 1:
        res = [fn(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(fn, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    return {"value": cores * m / busy, "unit": "configs/s", "cores": cores,
            "kind": "reference" if use_ref else "port",
            "sample": "%d S-%s samples per core (seed 1001), compute time of the slowest core %.2f s, wall %.2f s; %s" % (
                m, wl_eff.upper(), busy, wall,
                "the reference's own compiled code (tests/trbspcs via oracle/_ref)" if use_ref else
                "oracle/ C++ restatement of the reference path (real KDL/PQP sources are not in the reference tree)"),
            "n_valid": sum(r[1] for r in res)}


def run_reference(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    wl = args.workload
    cores = os.cpu_count() or 1
    n = args.samples or (100000 if wl == "rbs" else 1_000_000)
    K, W = args.steps, args.warmup
    budget = max(2.0, min(12.0, 150.0 / max(K + W, 1)))          # whole run stays within a few minutes
    W_run = min(W, 1)             # one 1-second warm-up pass at most (page-in of the binaries / meshes); printed as such
    for _ in range(W_run):
        cpu_baseline(wl, cores, 1.0)
    t0 = time.perf_counter()
    vals = []
    base = None
    for k in range(K):
        base = cpu_baseline(wl, cores, budget)
        vals.append(base["value"])
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    base["value"] = value
    unit = "edges/s" if wl == "rbs" else "configs/s"
    line = {"impl": "reference", "metric": "RBS edges/s (checkMotionRBS, PCS)" if wl == "rbs" else "validated closed-chain configs/s (project + FK/limits + collide)", "value": value, "unit": unit,
            "n_gpus": args.gpus, "steps": K, "warmup": W_run, "warmup_requested": W, "ms_per_step": 1e3 * wall / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "same generator and workload as the native arm (%s), each step a bounded sample sized for ~%.0f s on %d host cores" % (wl, budget, cores),
                       "units_per_gpu_per_step": n},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
