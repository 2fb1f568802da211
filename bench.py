#!/usr/bin/env python
"""bench.py -- encode+decode throughput of the ICER hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c2noise|c3|c4|c5|all]

A "step" is one pass of the hot path over one batch of synthetic frames: batch ENCODE (forward DWT -> LL mean /
sign-magnitude -> bit-plane context model + interleaved entropy coder -> segment assembly) followed by batch DECODE
of the streams just produced (scan -> entropy decoder -> inverse DWT).  The metric is
    Mpixels/s = frames * W * H / (t_encode + t_decode) / 1e6     (whole job, all ranks).
`value`  : inputs already resident in HBM, outputs left in HBM (device pointers through the C ABI).
`e2e`    : the same step through the C ABI with HOST (pinned) buffers; H2D of the frames, D2H of the streams,
           H2D of the streams and D2H of the decoded frames are all inside the timed region.  Two host threads with one
           codec context each work on alternating steps (the C ABI's documented double-buffering pattern):
           the copies of one step overlap the kernels of the other.
N > 1    : one process per GPU (torchrun), the batch is sharded by frame index -- frames are independent, so there
           is no data-path collective (weak scaling: every rank encodes+decodes its own `batch` frames); time is the
           max over ranks of the device-timed region.  `--single-process` instead drives all N GPUs from ONE process
           through icer_b200_*_batch_multi (the C ABI's own multi-GPU entry point).

The line's top-level keys describe BASELINE.json configs[1] (`c2`, the configuration the metric is quoted on).  The default
run then measures the other BASELINE configurations with a bounded number of steps and reports each of them -- value, e2e,
roofline, cpu_baseline, parity -- under `configs` (N = 1: c2noise, c3, c4, c5; N > 1: c5, the configuration BASELINE.json
shards across the GPUs of one box).

`--impl reference` times the UNMODIFIED reference (oracle/_ref/libicer_ref*.so, lib_icer compiled with its own
-O2 flags) on the box's host cores, one forked worker per core because lib_icer is non-re-entrant.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "c2": dict(desc="batch=1024 512x512 uint8-in-uint16 gray (natural-like synthetic), 6 stages, 10 segments, 9 bit planes, filter A, quota=W*H",
               batch=1024, w=512, h=512, chan=1, stages=6, segments=10, planes=9, max_stages=6, quota=512 * 512, kind=0, seed=20260922 + 2),
    "c2noise": dict(desc="batch=1024 512x512 uint8 noise in uint16 gray, 6 stages, 10 segments, 9 bit planes, filter A, quota=2*W*H",
                    batch=1024, w=512, h=512, chan=1, stages=6, segments=10, planes=9, max_stages=6, quota=2 * 512 * 512, kind=1, seed=20260922 + 2),
    "c2filtB": dict(desc="batch=1024 512x512 uint8-in-uint16 gray, 6 stages, 10 segments, 9 bit planes, filter B, quota=W*H",
                    batch=1024, w=512, h=512, chan=1, stages=6, segments=10, planes=9, max_stages=6, quota=512 * 512, kind=0, seed=20260922 + 2, filt=1),
    "c3": dict(desc="batch=1024 512x512 YUV planar uint16, 4 stages, 10 segments, 9 bit planes, filter A, quota=70000 B",
               batch=1024, w=512, h=512, chan=3, stages=4, segments=10, planes=9, max_stages=6, quota=70000, kind=0, seed=20260922 + 3),
    "c4": dict(desc="batch=256 4096x4096 12-bit-in-uint16 gray, 7 stages, 10 segments, 15 bit planes, filter A, quota=2*W*H",
               batch=256, w=4096, h=4096, chan=1, stages=7, segments=10, planes=15, max_stages=7, quota=2 * 4096 * 4096, kind=2, seed=20260922 + 4),
    "c5": dict(desc="batch=1024/GPU 1024x1024 uint8-in-uint16 gray, 6 stages, 10 segments, 9 bit planes, filter A, quota=W*H",
               batch=1024, w=1024, h=1024, chan=1, stages=6, segments=10, planes=9, max_stages=6, quota=1024 * 1024, kind=0, seed=20260922 + 5),
}
METRIC = "encode+decode Mpixels/s"


def usable_cores():
    """Host cores this process may really use: affinity mask, capped by the cgroup CPU quota (cpu.max)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def config_of(wl, world, frames_per_gpu):
    """The `config` object of a line: identical for our arm and the reference arm of the same workload."""
    n, w, h, chan = frames_per_gpu, wl["w"], wl["h"], wl["chan"]
    return {"workload": wl["desc"], "frames_per_gpu": n,
            "parallelism": "frames sharded across %d GPU(s), no data-path collective" % world,
            "l2": "inputs (%.0f MB/step) exceed the 126 MB L2; no flush needed" % (n * chan * w * h * 2 / 1e6),
            "timing": "CUDA events on the launching stream, max over ranks"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


def gpu_numa_cpus(index):
    """CPUs of the NUMA node the GPU hangs off (pinned staging buffers should be allocated and touched from there)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------
# CPU reference leg (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wl, frames_idx, planes15 = args
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    ref = harness.Ref(planes15)
    orc = harness.Oracle()
    imgs = []
    for f in frames_idx:
        if wl["chan"] == 1:
            imgs.append([orc.synth(wl["w"], wl["h"], f, wl["seed"], wl["kind"])])
        else:
            imgs.append([orc.synth(wl["w"], wl["h"], f + 1000 * c, wl["seed"], wl["kind"]) for c in range(3)])
    t_enc = t_dec = 0.0
    out = []
    for f, im in zip(frames_idx, imgs):
        t0 = time.perf_counter()
        enc = ref.compress(im, wl["stages"], wl.get("filt", 0), wl["segments"], wl["quota"])
        t1 = time.perf_counter()
        dec = ref.decompress(enc["stream"], wl["chan"], wl["w"] * wl["h"], wl["stages"], wl.get("filt", 0), wl["segments"])
        t2 = time.perf_counter()
        t_enc += t1 - t0; t_dec += t2 - t1
        out.append((f, enc["rc"], enc["stream"], dec["rc"], b"".join(i.tobytes() for i in dec["images"]) if dec["images"] is not None else None))
    return t_enc, t_dec, out


def cpu_reference_run(wl, frames_per_core, cores=None, keep=False):
    """Run the compiled reference on `cores` forked workers, `frames_per_core` frames each.  Returns Mpix/s (all cores)."""
    import multiprocessing as mp
    cores = cores or usable_cores()
    planes15 = wl["planes"] == 15
    jobs = [(wl, list(range(k * frames_per_core, (k + 1) * frames_per_core)), planes15) for k in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    px = cores * frames_per_core * wl["w"] * wl["h"]
    worst = max(r[0] + r[1] for r in res)
    one_core = sum(r[0] + r[1] for r in res) / cores
    outs = [o for r in res for o in r[2]] if keep else None
    return dict(mpix_s=px / worst / 1e6, mpix_s_1core=frames_per_core * wl["w"] * wl["h"] / one_core / 1e6, cores=cores,
                frames=cores * frames_per_core, wall_s=wall, enc_s=sum(r[0] for r in res) / cores, dec_s=sum(r[1] for r in res) / cores,
                outs=outs)


def frames_per_core_for(wl, target_s=6.0):
    # ~0.35 us / pixel / plane-pass for encode+decode on one 2 GHz core (512^2 -> ~0.18 s, 4096^2x15 -> ~21 s)
    est = wl["w"] * wl["h"] * wl["chan"] * wl["planes"] / 9.0 * 0.7e-6
    return max(1, int(target_s / est))


def reference_line(args, wl, steps, warmup):
    fpc = frames_per_core_for(wl)
    if os.environ.get("ICER_BENCH_FRAMES_PER_CORE"):       # test hook: a shorter sample (tests/test_bench_contract_cpu.py)
        fpc = max(1, int(os.environ["ICER_BENCH_FRAMES_PER_CORE"]))
    vals = []
    for it in range(warmup + steps):
        r = cpu_reference_run(wl, fpc)
        if it >= warmup:
            vals.append(r)
    v = float(np.mean([r["mpix_s"] for r in vals]))
    px = vals[0]["frames"] * wl["w"] * wl["h"]
    return {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "Mpixels/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * px / (v * 1e6), 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": config_of(wl, args.gpus, wl["batch"]),
            "cpu_baseline": {"value": round(v, 3), "unit": "Mpixels/s", "cores": vals[0]["cores"], "kind": "reference",
                             "sample": "%d frames of the workload per step (%d per core), one forked lib_icer worker per core (oracle/_ref)"
                                       % (vals[0]["frames"], fpc),
                             "one_core_mpix_s": round(float(np.mean([r["mpix_s_1core"] for r in vals])), 3)},
            "e2e": {"value": round(v, 3), "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def run_reference_arm(args, names, rank):
    if rank != 0:
        return
    line = reference_line(args, dict(WORKLOADS[names[0]], **batch_override(args, names[0])), args.steps, args.warmup)
    rest = []
    for nm in names[1:]:
        r = reference_line(args, WORKLOADS[nm], 1, 0)
        r["name"] = nm
        rest.append(r)
    if rest:
        line["configs"] = rest
    print(json.dumps(line), flush=True)


def batch_override(args, name):
    if not args.batch:
        return {}
    wl = WORKLOADS[name]
    return {"batch": args.batch, "desc": wl["desc"].replace("batch=%d" % wl["batch"], "batch=%d" % args.batch)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class Bench:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import icer_compression_b200 as icer
        self.torch, self.icer, self.args = torch, icer, args
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.devices = list(range(args.gpus)) if args.single_process else [local_rank]
        torch.cuda.set_device(self.devices[0])
        self.dev = torch.device("cuda", self.devices[0])
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.ngpu = world * len(self.devices)

    def barrier(self):
        for d in self.devices:
            self.torch.cuda.synchronize(d)
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        from icer_compression_b200.shard import max_over_ranks
        return max_over_ranks(vals, device=self.dev)

    # ---- one workload ---------------------------------------------------------------------------------------------
    def run(self, name, wl, steps, warmup, e2e_steps, with_cpu):
        torch, icer = self.torch, self.icer
        G = len(self.devices)
        n_gpu, w, h, chan, quota = wl["batch"], wl["w"], wl["h"], wl["chan"], wl["quota"]
        n = n_gpu * G                                   # frames this process codes per step
        params = (w, h, chan, wl["stages"], wl.get("filt", 0), wl["segments"], quota)
        multi = icer.IcerMulti(self.devices, planes=wl["planes"], max_stages=wl["max_stages"])
        ctx = multi.ctxs[0]
        dev = self.dev
        st = torch.cuda.current_stream().cuda_stream
        first = self.rank * n

        frames = torch.empty((n, chan, h, w), dtype=torch.int16, device=dev)
        if chan == 1:
            ctx.synth_frames(frames, n, w, h, first, wl["seed"], wl["kind"], st)
        else:
            tmp = torch.empty((n, h, w), dtype=torch.int16, device=dev)
            for c in range(3):   # Y,U,V = generator frames f, f+1000, f+2000
                ctx.synth_frames(tmp, n, w, h, first + 1000 * c, wl["seed"], wl["kind"], st)
                frames[:, c] = tmp
            del tmp
        out = torch.empty((n, quota), dtype=torch.uint8, device=dev)
        used = torch.zeros(n, dtype=torch.int64, device=dev)
        rc = torch.zeros(n, dtype=torch.int32, device=dev)
        dec = torch.empty((n, chan, h * w), dtype=torch.int16, device=dev)
        dims = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        drc = torch.zeros(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()

        # every call goes through the library's multi-context entry points (one context per GPU of this process)
        def encode():
            multi.compress_batch_raw(params, frames, n, out, quota, used, rc, icer.MEM_DEVICE)

        def decode():
            multi.decompress_batch_raw(params, out, quota, used, n, dec, dims, drc, icer.MEM_DEVICE)

        for c in multi.ctxs:
            c.enable_timing(True)
        for _ in range(warmup):
            encode(); decode()
        self.barrier()
        launches0 = multi.launch_count()
        sampler = ClockSampler(self.devices[0])
        if self.rank == 0:
            sampler.start()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
        kern_enc_ms, kern_dec_ms, enc_dwt_ms, dec_dwt_ms = [], [], [], []
        ev[0].record()
        for k in range(steps):
            encode()
            ev[2 * k + 1].record()
            kern_enc_ms.append(ctx.phase_ms(2)); enc_dwt_ms.append(ctx.phase_ms(1))
            decode()
            ev[2 * k + 2].record()
            kern_dec_ms.append(ctx.phase_ms(2)); dec_dwt_ms.append(ctx.phase_ms(1))
        self.barrier()
        clocks = sampler.stop() if self.rank == 0 else None
        launches = multi.launch_count() - launches0
        total_ms = ev[0].elapsed_time(ev[-1])
        enc_ms = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps)) / steps
        dec_ms = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(steps)) / steps
        total_ms, enc_ms, dec_ms = self.max_over_ranks([total_ms, enc_ms, dec_ms])   # the slowest rank defines the step
        ms_per_step = total_ms / steps
        px_step = self.world * n * w * h
        value = px_step / (ms_per_step * 1e-3) / 1e6
        for c in multi.ctxs:
            c.enable_timing(False)

        used_h = used.cpu().numpy()
        rc_h = rc.cpu().numpy()
        drc_h = drc.cpu().numpy()
        stream_bytes = int(used_h.sum())
        ok_frames = int(((rc_h == 0) | (rc_h == -5)).sum())
        roundtrip = None
        # 8-bit frames with all 9 planes decode to the input.  (12-bit frames do not, in lib_icer either: the LL mean travels through an
        # 8-bit header field, icer.h:270 -- the decode is compared with the reference's own decode instead, see `parity`.)
        if (rc_h == 0).all() and wl["kind"] == 0 and wl["planes"] >= 9:
            roundtrip = bool(torch.equal(dec.view(n, chan, h, w), frames))

        # ---- CPU baseline on a bounded sample of the same workload + bit-exactness of that sample (rank 0) -----------------
        cpu_baseline = parity = None
        if with_cpu and self.rank == 0:
            fpc = frames_per_core_for(wl, 5.0)
            cores = usable_cores()
            while cores * fpc > n and fpc > 1:
                fpc -= 1
            cores = min(cores, n)
            r = cpu_reference_run(wl, fpc, cores=cores, keep=True)
            idx = sorted(o[0] for o in r["outs"])
            out_h = out[idx].cpu().numpy()
            dec_h = dec[idx].cpu().numpy()
            bad = 0
            for f, rrc, rstream, rdrc, rdec in r["outs"]:     # encode status + stream bytes, decode status + the REFERENCE'S decoded samples
                j = idx.index(f)
                ln = int(used_h[f])
                if rrc != int(rc_h[f]) or bytes(out_h[j, :ln]) != rstream or rdrc != int(drc_h[f]) or (rdec is not None and dec_h[j].tobytes() != rdec):
                    bad += 1
            cpu_baseline = {"value": round(r["mpix_s"], 3), "unit": "Mpixels/s", "cores": r["cores"], "kind": "reference",
                            "sample": "%d frames of the workload (%d per core), lib_icer compiled -O2 (oracle/_ref), one forked worker per core"
                                      % (r["frames"], fpc), "one_core_mpix_s": round(r["mpix_s_1core"], 3)}
            parity = {"frames_checked": len(r["outs"]), "mismatches": bad, "against": "oracle/_ref (unmodified lib_icer): encode rc + stream "
                      "bytes, decode rc + decoded samples"}
        if self.dist is not None:
            self.dist.barrier()

        # ---- end to end through the C ABI with host buffers ---------------------------------------------------------------
        e2e = None
        if e2e_steps > 0:
            e2e = self.run_e2e(name, wl, multi, frames, params, e2e_steps, px_step)
            if isinstance(e2e, dict) and "error" in e2e:
                sys.stderr.write("[bench] e2e of %s failed: %s\n" % (name, e2e["error"]))
        multi.close()

        peak, peak_src = measured_peak_gbs()
        alg_bytes = n * chan * w * h * 2 + stream_bytes                 # per process = per launch set
        k4_ms, k7_ms = float(np.mean(kern_enc_ms)), float(np.mean(kern_dec_ms))
        step_bytes = 2 * alg_bytes                                       # encode: in + stream ; decode: stream + out
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(name, {})
            except Exception:
                traffic = {}
        alg1 = alg_bytes // G                                            # one GPU's launch

        def roof(kname, desc, ms):
            ach = alg1 / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            return {"bound": "hbm", "kernel": desc, "achieved": round(ach, 2), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 5),
                    "traffic": traffic.get(kname), "peak_source": peak_src, "algorithmic_bytes_per_launch": alg1, "kernel_ms": round(ms, 3)}

        r_enc = roof("k_encode", "k_encode (bit-plane context model + interleaved entropy coder)", k4_ms)
        r_dec = roof("k_decode", "k_decode (bit-plane context model + interleaved entropy decoder)", k7_ms)
        roofline, roofline_other = (r_dec, r_enc) if k7_ms >= k4_ms else (r_enc, r_dec)
        roofline["encode_dwt_ms"] = round(float(np.mean(enc_dwt_ms)), 3)
        roofline["decode_dwt_ms"] = round(float(np.mean(dec_dwt_ms)), 3)
        roofline["step_frac"] = round(step_bytes / G / (ms_per_step * 1e-3) / 1e9 / peak, 5)
        roofline["note"] = "instruction-issue bound integer kernels: the HBM fraction is reported as the contract asks, it is not the limiter"
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "Mpixels/s", "n_gpus": self.ngpu, "steps": steps,
            "warmup": warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": config_of(wl, self.ngpu, n_gpu),
            "encode_mpix_s": round(px_step / (enc_ms * 1e-3) / 1e6, 2), "decode_mpix_s": round(px_step / (dec_ms * 1e-3) / 1e6, 2),
            "encode_ms": round(enc_ms, 3), "decode_ms": round(dec_ms, 3),
            # a planar Y,U,V frame counts w*h pixels but 3*w*h samples: both are reported
            "samples_per_pixel": chan, "value_msamples_s": round(value * chan, 2),
            "compressed_bytes_per_step": stream_bytes, "frames_ok": ok_frames, "decode_rc_ok": int((drc_h == 0).sum()),
            "roundtrip_lossless": roundtrip,
            "roofline": roofline, "roofline_other": roofline_other,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if self.args.single_process:
            line["config"]["parallelism"] = "ONE process drives %d GPU(s) through icer_b200_*_batch_multi (frames sharded by index, peer scatter/gather from device 0)" % G
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
            line["parity"] = parity
        del frames, out, dec
        torch.cuda.empty_cache()
        return line

    # ---- e2e: host buffers, two workers on alternating steps ----------------------------------------------------------
    def run_e2e(self, name, wl, multi, frames, params, ksteps, px_step):
        torch, icer = self.torch, self.icer
        n, chan, h, w = frames.shape
        quota = wl["quota"]
        frame_bytes = n * chan * h * w * 2
        big = frame_bytes > (3 << 30)
        nworkers = 1 if (big or len(self.devices) > 1 or self.args.e2e_workers < 2) else min(self.args.e2e_workers, 4)
        if nworkers > 2 and frame_bytes > (3 << 29):
            nworkers = 2            # every worker pins its own output buffers: keep the pinned host memory of an 8-rank job bounded
        old_aff = os.sched_getaffinity(0)
        cpus = gpu_numa_cpus(self.devices[0]) if len(self.devices) == 1 else None
        if cpus:
            os.sched_setaffinity(0, cpus)          # pinned staging memory gets allocated and first-touched on the GPU's NUMA node
        try:
            fh = torch.empty((n, chan, h, w), dtype=torch.int16).pin_memory()
            fh.copy_(frames)
            workers = []
            for k in range(nworkers):
                m = multi if k == 0 else icer.IcerMulti(self.devices, planes=wl["planes"], max_stages=wl["max_stages"])
                workers.append(dict(m=m, oh=torch.empty((n, quota), dtype=torch.uint8).pin_memory(), uh=torch.zeros(n, dtype=torch.int64),
                                    rh=torch.zeros(n, dtype=torch.int32), dh=torch.empty((n, chan, h * w), dtype=torch.int16).pin_memory(),
                                    dimh=torch.zeros((n, 2), dtype=torch.int32), drh=torch.zeros(n, dtype=torch.int32)))

            def step_host(wk):
                wk["m"].compress_batch_raw(params, fh, n, wk["oh"], quota, wk["uh"], wk["rh"], icer.MEM_HOST)
                wk["m"].decompress_batch_raw(params, wk["oh"], quota, wk["uh"], n, wk["dh"], wk["dimh"], wk["drh"], icer.MEM_HOST)

            def run_steps(count):
                # worker k takes steps k, k + nworkers, ...; the ctypes calls release the GIL
                def loop(k):
                    for _ in range(k, count, nworkers):
                        step_host(workers[k])
                th = [threading.Thread(target=loop, args=(k,)) for k in range(1, nworkers)]
                for t in th:
                    t.start()
                loop(0)
                for t in th:
                    t.join()

            run_steps(max(nworkers, min(self.args.warmup, 2)))
            self.barrier()
            # enough steps for the workers to fall out of step with each other (they all start on the same phase): the copies of one
            # worker then overlap the kernels of the others, which is the steady state a caller with a queue of batches sees
            ksteps = max(ksteps, nworkers) if nworkers == 1 else max(ksteps, 4 * nworkers)
            t0 = time.perf_counter()
            run_steps(ksteps)
            for d in self.devices:
                torch.cuda.synchronize(d)
            e2e_ms = (time.perf_counter() - t0) * 1e3 / ksteps
            e2e_ms = self.max_over_ranks([e2e_ms])[0]
            ok = all(bool(torch.equal(wk["dh"], workers[0]["dh"])) and bool(torch.equal(wk["uh"], workers[0]["uh"])) for wk in workers[1:])
            sb = int(workers[0]["uh"].sum().item())
            res = {"value": round(px_step / (e2e_ms * 1e-3) / 1e6, 2), "unit": "Mpixels/s", "ms_per_step": round(e2e_ms, 3), "steps": ksteps,
                   "h2d_bytes_per_step": int(n * chan * h * w * 2 + sb), "d2h_bytes_per_step": int(sb + n * chan * h * w * 2),
                   "host_workers": nworkers, "workers_agree": ok,
                   "note": "host pinned frames -> C ABI -> host streams -> C ABI -> host frames, wall clock over all steps; %s; only the used bytes "
                           "of every stream slot cross PCIe" % ("%d host threads with one context each work on alternating steps, so the copies of one "
                           "step overlap the kernels of the other" % nworkers if nworkers > 1 else "one host thread")}
            for wk in workers[1:]:
                wk["m"].close()
            return res
        finally:
            os.sched_setaffinity(0, old_aff)


def latency_section(args):
    """BASELINE config 1 (ONE 512x512 image through the reference's own single-image calls, host buffers) and small batches of the
    largest frames: what a caller without a big batch sees.  lane = frame needs 32 frames per warp; a single image runs with one lane
    per warp and its latency is the longest stream's (lib_icer needs 84 + 95 ms for the same image on one core)."""
    import statistics
    import numpy as np
    import torch
    import icer_compression_b200 as icer
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    out = {}
    boat = harness.load_boat()
    leg = icer.LegacyAPI()
    enc_ms, dec_ms = [], []
    stream = None
    for k in range(8):
        t0 = time.perf_counter()
        r = leg.compress([boat], 6, 0, 10, 512 * 512)
        t1 = time.perf_counter()
        d = leg.decompress(r["stream"], 1, 512 * 512, 6, 0, 10)
        t2 = time.perf_counter()
        if k >= 2:
            enc_ms.append((t1 - t0) * 1e3); dec_ms.append((t2 - t1) * 1e3)
        stream = r["stream"]
    c1 = {"workload": "single 512x512 uint8 grayscale (boat.512), 6 stages, 10 segments, 9 bit planes, icer_compress_image_uint16 / icer_decompress_image_uint16 "
                      "with host buffers, wall clock incl. the Python wrapper's buffer copies, median of 6 calls",
          "encode_ms": round(statistics.median(enc_ms), 3), "decode_ms": round(statistics.median(dec_ms), 3),
          "lossless": bool(d["rc"] == 0 and (d["images"][0] == boat).all()), "stream_bytes": len(stream)}
    if not args.no_cpu and harness.have_ref():
        ref = harness.Ref()
        t0 = time.perf_counter(); rr = ref.compress([boat], 6, 0, 10, 512 * 512); t1 = time.perf_counter()
        ref.decompress(rr["stream"], 1, 512 * 512, 6, 0, 10); t2 = time.perf_counter()
        c1["reference_encode_ms"] = round((t1 - t0) * 1e3, 3); c1["reference_decode_ms"] = round((t2 - t1) * 1e3, 3)
        c1["same_stream_as_reference"] = bool(rr["stream"] == stream)
    out["c1_single_image"] = c1
    # small batches of config 4's frames, device-resident
    wl = WORKLOADS["c4"]
    dev = torch.device("cuda", torch.cuda.current_device())
    w, h, quota = wl["w"], wl["h"], wl["quota"]
    for nb in (8, 32):
        ctx = icer.IcerB200(device=torch.cuda.current_device(), planes=wl["planes"], max_stages=wl["max_stages"])
        st = torch.cuda.current_stream().cuda_stream
        frames = torch.empty((nb, 1, h, w), dtype=torch.int16, device=dev)
        ctx.synth_frames(frames, nb, w, h, 0, wl["seed"], wl["kind"], st)
        o = torch.empty((nb, quota), dtype=torch.uint8, device=dev)
        used = torch.zeros(nb, dtype=torch.int64, device=dev); rc = torch.zeros(nb, dtype=torch.int32, device=dev)
        dec = torch.empty((nb, 1, h * w), dtype=torch.int16, device=dev)
        dims = torch.zeros((nb, 2), dtype=torch.int32, device=dev); drc = torch.zeros(nb, dtype=torch.int32, device=dev)
        params = (w, h, 1, wl["stages"], 0, wl["segments"], quota)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        for k in range(2):
            ev[0].record()
            ctx.compress_batch_raw(params, frames, nb, o, quota, None, None, used, rc, icer.MEM_DEVICE)
            ev[1].record()
            ctx.decompress_batch_raw(params, o, quota, used, nb, dec, dims, drc, icer.MEM_DEVICE)
            ev[2].record()
            torch.cuda.synchronize()
        e_ms, d_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
        out["c4_batch%d" % nb] = {"workload": "batch=%d 4096x4096 12-bit-in-uint16, 7 stages, 15 bit planes, device-resident" % nb,
                                  "encode_ms": round(e_ms, 2), "decode_ms": round(d_ms, 2),
                                  "value": round(nb * w * h / ((e_ms + d_ms) * 1e-3) / 1e6, 2), "unit": "Mpixels/s",
                                  "rc_ok": int(((rc == 0) | (rc == -5)).sum().item())}
        del frames, o, dec
        ctx.close()
        torch.cuda.empty_cache()
    return out


def run_ours(args, names, rank, world, local_rank):
    b = Bench(args, rank, world, local_rank)
    head = dict(WORKLOADS[names[0]], **batch_override(args, names[0]))
    line = b.run(names[0], head, args.steps, args.warmup, 0 if args.no_e2e else max(2, min(args.steps, 6)), not args.no_cpu)
    rest = []
    for nm in names[1:]:
        wl = WORKLOADS[nm]
        try:
            r = b.run(nm, wl, max(1, min(args.steps, 3)), max(1, min(args.warmup, 3)), 0 if args.no_e2e else 2, not args.no_cpu)
        except Exception as e:       # the headline line must not be lost to a failure of an additional configuration
            if b.dist is not None:
                raise                # ranks have to stay in step: a multi-rank job fails as a whole
            r = {"error": repr(e)}
        r["name"] = nm
        rest.append(r)
    if rank == 0:
        if rest:
            line["configs"] = rest
        if args.workload in (None, "all") and world == 1 and not args.single_process:
            try:
                line["latency"] = latency_section(args)
            except Exception as e:      # the headline must not depend on the extra measurements
                line["latency"] = {"error": repr(e)}
        print(json.dumps(line), flush=True)
    if b.dist is not None:
        b.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="one of %s, or 'all'; default: c2 as the headline + the other BASELINE configs under `configs`" % ", ".join(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="override frames per GPU of the headline workload")
    ap.add_argument("--single-process", action="store_true", help="drive all --gpus devices from this one process through icer_b200_*_batch_multi")
    ap.add_argument("--e2e-workers", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--watchdog", type=float, default=1500.0, help="dump every Python stack and exit if the run takes longer than this many seconds")
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog, exit=True)     # a stalled run must not hold a GPU box until the outer limit
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in (None, "all"):
        # the headline + the rest of BASELINE.json's configurations: everything that runs on one GPU at N = 1, config 5 at N > 1
        names = ["c2", "c2noise", "c3", "c4", "c5"] if args.gpus == 1 else ["c2", "c5"]
    else:
        names = [args.workload]
    if args.impl == "reference":
        if args.workload is None:
            names = ["c2"] if args.gpus == 1 else ["c2", "c5"]      # the reference arm of the headline (+ config 5 when the job is sharded)
        run_reference_arm(args, names, rank)
    else:
        run_ours(args, names, rank, world, local_rank)


if __name__ == "__main__":
    main()
