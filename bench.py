#!/usr/bin/env python
"""bench.py -- ME+TQ frames/s of the frame-level pre-pass (BASELINE.json metric) on N B200s.

A step = one pass of the hot path over one frame pair: motion search of every PU 64/32/16/8 (parent -> child chain,
integer walk + half + quarter pel), motion compensation luma + chroma at each size, and the inter T/Q chain over TU
32/32/16/8/4 (+ chroma) with reconstruction.  Workload: synthetic YUV 4:2:0 of SURVEY.md 8(d), IPPP quarter-pel, fixed QP 32.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 720p|1080p|2160p] [--impl ours|reference]

N > 1 is launched by torchrun (one rank per GPU); every rank processes its own independent stream of frames
(GOP-per-GPU batching, BASELINE.json configs[4]; no data-path collective), value = frames of all ranks / max time.
`--impl reference` times the reference's own CPU functions (oracle/_ref, compiled from /root/reference) over the same
pre-pass on the host cores of the box; rank 0 only.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {"720p": (1280, 720), "1080p": (1920, 1080), "2160p": (3840, 2160)}
QP, AVG_DIST = 32, 650.0
N_RESIDENT = 16           # distinct frames cycled through so that no step finds its inputs in L2
L2_BYTES = 126 * 2 ** 20


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes(pp, w, h):
    """per-launch algorithmic HBM bytes of every pre-pass kernel: each input plane read once,
    each output written once, for the PUs / TUs that launch covers."""
    out = {}
    luma = w * h
    for d in range(4):
        s = 64 >> d
        n_pu = (w // s) * (h // s)
        out[f"me{s}"] = 2 * luma + 24 * n_pu                    # cur + ref luma (u8) in, one result record per PU out
        out[f"mc{s}"] = int(1.5 * n_pu * s * s) * 2 + 8 * n_pu  # ref in, pred out (luma + chroma), mv in
    for p in range(5):
        for c in range(3):
            t = pp.tu_size(p, c)
            if not t:
                continue
            n = pp.num_tus(p, c)
            out[f"tq{p}{'yuv'[c]}{t}"] = n * t * t * (1 + 1 + 1 + 2) + n * (8 + 16)   # cur, pred in; recon, levels out; job, result
    return out


def run_reference(args, w, h, rank, world):
    """the reference's own CPU functions over the same pre-pass, all host threads; a step = `sample` frames"""
    from homerhevc_b200 import synth
    from _oracle import have_ref, ref_prepass
    if rank != 0:
        return
    if not have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not in this tree"}))
        return
    cores = len(os.sched_getaffinity(0))
    tex = synth.make_texture(w, h)
    frames = [synth.make_frame(tex, w, h, n) for n in range(5)]
    sample = 1 if (w, h) == WORKLOADS["2160p"] else (2 if (w, h) == WORKLOADS["1080p"] else 4)
    def step(i):
        t = 0.0
        for k in range(sample):
            j = (i * sample + k) % 4
            s, _ = ref_prepass(frames[j + 1], frames[j], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False)
            t += s
        return t
    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    secs = time.perf_counter() - t0
    fps = args.steps * sample / secs
    line = {"impl": "reference", "metric": "ME+TQ frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int16/int32", "data": "synthetic",
            "config": {"workload": workload_name(args, w, h), "frames_per_step": sample, "qp": QP, "avg_dist": AVG_DIST},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample} frame(s) per step, reference functions via oracle/_ref, {cores} threads, CTUs dealt round-robin"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    # for context (not the metric): the WHOLE unmodified reference encoder -- mode decision, CABAC, in-loop filters included -- on
    # the same kind of input, one engine, WPP off, fixed QP (the configuration whose output the parity tests pin)
    try:
        import ctypes as C
        import numpy as np
        from _oracle import ref
        _, D = ref()
        nf = 3
        clip = synth.make_clip(w, h, nf, seed=5)
        yuv = np.concatenate([np.concatenate([p.reshape(-1) for p in f]) for f in clip])
        bs = np.zeros(16 << 20, np.uint8); enc_secs = C.c_double(0)
        n = D.refdrv_encode_lockstep(w, h, nf, yuv.ctypes.data_as(C.POINTER(C.c_uint8)), QP, 1, 0, -1, bs.ctypes.data_as(C.POINTER(C.c_uint8)), bs.size,
                                     None, None, None, C.byref(enc_secs))
        if n > 0 and enc_secs.value > 0:
            line["whole_encoder"] = {"value": nf / enc_secs.value, "unit": "frames/s", "frames": nf, "bytes": int(n),
                                     "what": "unmodified reference encoder (homer_lib), IPP, 1 engine, WPP off, fixed QP, all stages"}
    except Exception as e:
        line["whole_encoder"] = {"value": None, "what": f"failed: {e}"}
    print(json.dumps(line))


def run_intra(args, w, h, rank, world, local, hb, synth):
    """the intra mode pre-search of one picture -- SADs of all 35 modes for every 32/16/8/4 luma block, reference
    samples from the original picture -- through hb_intra_run (host job list and samples in, host SAD table out, copies inside
    the timed call), next to the reference's own predictors + sad on all host cores.  Rank 0 only; not the headline metric."""
    if rank != 0:
        return
    from homerhevc_b200.intra_jobs import presearch_jobs
    ctx = hb.Context(local)
    tex = synth.make_texture(w, h)
    frames = [synth.make_frame(tex, w, h, n) for n in range(4)]
    from homerhevc_b200.lib import presearch_records
    prep = [presearch_jobs(f[0]) for f in frames]
    rec = presearch_records(prep[0][0])                     # the block list is the same for every frame of this size
    # samples and SAD table in pinned host memory, as an encoder would keep them
    adi_pin = [ctx.pinned(p[1].nbytes).view(np.int16) for p in prep]
    for a, p in zip(adi_pin, prep):
        a[:] = p[1]
    sad_pin = ctx.pinned(len(rec) * 35 * 4).view(np.uint32).reshape(len(rec), 35)
    dev = [hb.Frame(ctx, w, h) for _ in frames]
    for d, f in zip(dev, frames):
        d.upload_u8(*f)
    ctx.sync()
    for i in range(max(3, args.warmup)):
        ctx.intra_presearch(dev[i % 4], rec, adi_pin[i % 4], sad_pin)
    steps = min(args.steps, 40)
    t0 = time.perf_counter()
    for i in range(steps):
        sads = ctx.intra_presearch(dev[i % 4], rec, adi_pin[i % 4], sad_pin)
    secs = time.perf_counter() - t0
    jobs, adi, off = prep[(steps - 1) % 4]
    line = {"metric": "intra 35-mode pre-search frames/s", "value": steps / secs, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int32 accumulate",
            "data": "synthetic", "config": {"workload": workload_name(args, w, h) + ", every 32/16/8/4 luma block, reference samples from the original picture",
                                            "blocks": int(len(jobs)), "modes": 35, "timing": "wall clock around the blocking API call, host buffers in and out"},
            "e2e": {"value": steps / secs, "unit": "frames/s", "h2d_bytes_per_step": int(jobs.nbytes * 2 + adi.nbytes), "d2h_bytes_per_step": int(sads.nbytes)},
            "gpu_launches": steps}
    try:
        from _oracle import have_ref, ref_intra_presearch
        if have_ref() and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            ref_secs, ref_sads = ref_intra_presearch(frames[(steps - 1) % 4][0], jobs, adi, off, n_threads=cores)
            line["cpu_baseline"] = {"value": 1.0 / ref_secs, "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "1 frame, the reference's create_intra_*_prediction + adi_filter + sad (oracle/_ref)",
                                    "identical_to_gpu": bool((ref_sads == sads).all())}
    except Exception as e:
        line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))


def run_finalise(args, w, h, rank, world, local, hb, synth):
    """reference-frame finalisation of one picture -- deblocking (strengths derived on the device from per-unit mode
    data), SAO statistics, SAO offset pass, border -- through the blocking C API with host buffers, next to the reference's own
    functions (one host thread: the reference deblocks a picture on a single thread).  Rank 0 only; not the headline metric."""
    if rank != 0:
        return
    from _oracle import have_ref, random_deblock_case, random_sao_params, ref_deblock, ref_sao_apply, ref_sao_stats
    rng = np.random.default_rng(7)
    ctx = hb.Context(local)
    m, planes = random_deblock_case(rng, w, h)
    units = np.zeros(m["qp"].shape, hb.lib.UNIT_INFO_DT)
    units["cu_depth"], units["tu_depth"], units["intra"], units["cbf_luma"], units["qp"] = m["cu"], m["tu"], m["intra"], m["cbf"], m["qp"]
    units["ref_idx"] = np.where(m["intra"] != 0, -1, 0); units["mvx"] = m["mv"][..., 0]; units["mvy"] = m["mv"][..., 1]
    org = [np.clip(p.astype(np.int16) + rng.integers(-4, 5, p.shape), 0, 255).astype(np.uint8) for p in planes]
    types, offs = random_sao_params(rng, w, h)
    rec, fin, src = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
    src.upload_u8(*org)

    def step():
        rec.upload_u8(*planes)                                   # stands for the reconstruction the T/Q kernels left on the device
        ctx.deblock_units(rec, units, 2, 2)
        st = ctx.sao_stats(src, rec)
        ctx.sao_apply(rec, fin, types, offs)                     # the host's SAO decision would sit between the two calls
        return st
    for _ in range(max(3, args.warmup)):
        step()
    steps = min(args.steps, 50)
    t0 = time.perf_counter()
    for _ in range(steps):
        st = step()
    secs = time.perf_counter() - t0
    line = {"metric": "reference-frame finalisation frames/s", "value": steps / secs, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int32 accumulate", "data": "synthetic",
            "config": {"workload": f"{w}x{h}: deblocking (random CU/TU trees, modes, cbf, QPs, vectors) + SAO statistics + SAO offset pass + border",
                       "timing": "wall clock around the blocking API calls, host buffers in and out (the picture upload included)"},
            "e2e": {"value": steps / secs, "unit": "frames/s", "h2d_bytes_per_step": int(sum(p.nbytes for p in planes) + units.nbytes + types.nbytes + 2 * offs.nbytes // 4),
                    "d2h_bytes_per_step": int(st.nbytes + 2 * units.size)},
            "gpu_launches": int(steps * 9)}
    if have_ref() and not args.no_cpu_baseline:
        try:
            t0 = time.perf_counter()
            dexp, bsv, bsh, _ = ref_deblock(planes, w, h, m)
            rst = ref_sao_stats(dexp, org, w, h)
            rfin = ref_sao_apply(dexp, w, h, types, offs)
            ref_secs = time.perf_counter() - t0
            got = fin.download()
            line["cpu_baseline"] = {"value": 1.0 / ref_secs, "unit": "frames/s", "cores": 1, "kind": "reference",
                                    "sample": "1 frame: hmr_deblock_filter_cu per CTU and direction, get_sao_stats, sao_offset_ctu (oracle/_ref), one thread, int16 conversions included",
                                    "identical_to_gpu": bool(all((a == b).all() for a, b in zip(got, rfin)) and all((st[f] == rst[f]).all() for f in st.dtype.names))}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
    print(json.dumps(line))


def run_bands(args, w, h, rank, world, local, torch, dist, hb, synth, barrier):
    """strong scaling of ONE stream of frames: every GPU owns a CTU-row band; per frame the reference rows a band needs from
    its neighbours (68 luma / 36 chroma rows per side) are exchanged over NCCL, then the band's pre-pass runs"""
    from homerhevc_b200 import bands
    ctx = hb.Context(local)
    tex = synth.make_texture(w, h)
    n_pairs = 4
    host = [synth.make_frame(tex, w, h, n) for n in range(n_pairs + 1)]
    frames = [hb.Frame(ctx, w, h) for _ in range(n_pairs + 1)]
    for f, p in zip(frames, host):
        f.upload_u8(*p)
    ctu_rows = (h + 63) // 64
    row0, nrows = bands.band_ctu_rows(ctu_rows, world, rank)
    pp = hb.Prepass(ctx, w, h, qp=QP, use_graph=1, band=(row0, nrows))
    ex = bands.FrameHaloExchanger(torch, dist, ctx, w, h, world, rank, torch.device("cuda", local)) if world > 1 else None

    def step(i):
        j = i % n_pairs
        if ex:
            ex.exchange(frames[j])            # the reference of this frame: halos from the neighbours over NVLink
        pp.run(frames[j + 1], frames[j], AVG_DIST)

    for i in range(max(args.warmup, n_pairs)):
        step(i)
    ctx.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    ctx.sync()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    if rank == 0:
        print(json.dumps({"metric": "ME+TQ frames/s", "value": args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "u8 samples, int16 residual/levels, int32 accumulate", "data": "synthetic",
                          "config": {"workload": workload_name(args, w, h), "parallelism": f"ctu-row bands x{world}, NCCL halo exchange",
                                     "halo_rows": {"luma": bands.HALO_LUMA, "chroma": bands.HALO_CHROMA},
                                     "halo_bytes_sent_per_frame_rank0": ex.bytes_per_exchange if ex else 0,
                                     "timing": "wall clock between device-synchronised barriers (host drives the NCCL exchange)"}}))


def workload_name(args, w, h):
    return f"{w}x{h} IPPP quarter-pel ME (PU 64/32/16/8) + MC + inter T/Q (TU 32/32/16/8/4), fixed QP {QP}, synthetic YUV420"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=16, help="independent GOP streams in flight per GPU")
    ap.add_argument("--mode", default="gops", choices=["gops", "bands", "intra", "finalise"],
                    help="gops: independent GOP streams per GPU (default, weak scaling); bands: one frame split into CTU-row bands "
                         "across the GPUs with an NCCL halo exchange of the reference (BASELINE.json configs[3], strong scaling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w, h = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.steps > 20:
            args.steps, args.warmup = 5, 1
        run_reference(args, w, h, rank, world)
        return

    import torch
    import torch.distributed as dist
    import homerhevc_b200 as hb
    from homerhevc_b200 import synth

    if os.environ.get("HB_PIN_CPUS") == "1" and world > 1:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
    if os.environ.get("HB_BLOCKING") == "1":
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        cu.cuInit(0)
        print("blocking flags rc", cu.cuDevicePrimaryCtxSetFlags_v2(local, 4), file=sys.stderr)
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.mode == "finalise":
        run_finalise(args, w, h, rank, world, local, hb, synth)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.mode == "intra":
        run_intra(args, w, h, rank, world, local, hb, synth)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.mode == "bands":
        run_bands(args, w, h, rank, world, local, torch, dist, hb, synth, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    ctx = hb.Context(local)
    tex = synth.make_texture(w, h)
    # every rank works on its own stream of frames (its own GOP): same texture, different pan phase
    host = [synth.make_frame(tex, w, h, n + 5 * rank, seed=rank) for n in range(N_RESIDENT + 1)]
    frame_bytes = w * h * 3 // 2
    pin = ctx.pinned(frame_bytes * (N_RESIDENT + 1))
    pinned = []
    for i, (y, u, v) in enumerate(host):
        base = pin[i * frame_bytes:(i + 1) * frame_bytes]
        py = base[:w * h].reshape(h, w); pu = base[w * h:w * h * 5 // 4].reshape(h // 2, w // 2); pv = base[w * h * 5 // 4:].reshape(h // 2, w // 2)
        py[:], pu[:], pv[:] = y, u, v
        pinned.append((py, pu, pv))
    resident = [hb.Frame(ctx, w, h) for _ in range(N_RESIDENT + 1)]
    for f, p in zip(resident, pinned):
        f.upload_u8(*p)
    pp = hb.Prepass(ctx, w, h, qp=QP, use_graph=1)
    out_bytes = pp.output_bytes()
    out_pin = ctx.pinned(out_bytes)
    ctx.sync()

    # N_SLOTS independent streams of frames (GOPs, BASELINE.json configs[4]) are in flight per GPU: each has its own context
    # (CUDA stream), pre-pass plan and output buffers, so the search chain of one frame overlaps the T/Q tail of another
    N_SLOTS, LAMBDA = max(1, args.streams), 60
    slots = []
    for k in range(N_SLOTS):
        c = hb.Context(local)
        spp = hb.Prepass(c, w, h, qp=QP, use_graph=1, compact_tables=2)    # compact ME records + one cost record per coding unit on the wire (e2e); `value` never fetches
        n_ctus = spp.num_ctus()
        slots.append({"ctx": c, "cur": hb.Frame(c, w, h), "ref": hb.Frame(c, w, h), "pp": spp, "tables": c.pinned(spp.tables_bytes()),
                      "out": c.pinned(frame_bytes + 4 * w * h), "sel": np.zeros(n_ctus, np.uint8), "off": np.zeros(n_ctus + 1, np.int32), "d2h": 0})

    def step_resident(i):
        """one step = one frame on every in-flight stream, inputs resident in HBM"""
        for k, sl in enumerate(slots):
            j = (i + 4 * k) % N_RESIDENT
            sl["pp"].run(resident[j + 1], resident[j], AVG_DIST)

    def sync_all():
        for sl in slots:
            sl["ctx"].sync()
        ctx.sync()

    # ---- value: inputs resident in HBM, device time of exactly K steps (CUDA events on the library's stream; the
    # in-flight streams are ordered after the start event and before the stop event on the device)
    for i in range(N_RESIDENT):          # one-time CUDA graph capture per (stream, cur, ref), outside warm-up and timing
        step_resident(i)
    for i in range(args.warmup):
        step_resident(i)
    sync_all()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(sl["ctx"].launch_count() for sl in slots)
    ctx.timer_begin()
    for sl in slots:
        sl["ctx"].wait(ctx)
    for i in range(args.steps):
        step_resident(i)
    for sl in slots:
        ctx.wait(sl["ctx"])
    ms = ctx.timer_end()
    launches = sum(sl["ctx"].launch_count() for sl in slots) - l0
    barrier()
    clocks = sampler.stop()

    # ---- e2e: the call sequence a host encoder makes, with HOST buffers, per frame:
    #   upload cur + ref (pinned, async) -> pre-pass -> fetch the cost tables -> host picks a depth per CTU (hb_prepass_select,
    #   the stand-in for the host's mode decision) -> gather + fetch the reconstruction and coded levels of that choice.
    # N_SLOTS independent streams of frames (GOPs) are in flight per GPU so that copies, kernels and the host step overlap.
    def begin_frame(sl, i):
        j = i % N_RESIDENT
        sl["pp"].frame_begin(sl["cur"], sl["ref"], pinned[j + 1], pinned[j], AVG_DIST, sl["tables"])

    def finish_frame(sl):
        sl["d2h"] += sl["pp"].frame_finish(LAMBDA, sl["tables"], sl["sel"], sl["off"], sl["out"]) + sl["tables"].nbytes

    E2E_THREADS = int(os.environ.get("HB_E2E_THREADS", max(1, N_SLOTS // 2)))

    def run_e2e(n):
        # one host thread per TWO in-flight streams (the reference runs one pthread per encoder engine, hmr_encoder_lib.c:1647):
        # a thread queues frame n+1 on its second stream (hb_prepass_frame_begin returns at once) before it blocks in
        # hb_prepass_frame_finish of frame n.  The C calls release the GIL, so the threads overlap as well.
        def worker(k):
            mine = [slots[k], slots[k + E2E_THREADS]] if k + E2E_THREADS < N_SLOTS else [slots[k]]
            pending = None
            for c, i in enumerate(range(k, n, E2E_THREADS)):
                sl = mine[c % len(mine)]
                if pending is sl:                       # single-stream fallback: finish before reusing the stream
                    finish_frame(pending); pending = None
                begin_frame(sl, i)
                if pending is not None:
                    finish_frame(pending)
                pending = sl
            if pending is not None:
                finish_frame(pending)
        ths = [threading.Thread(target=worker, args=(k,)) for k in range(E2E_THREADS)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    # long enough for a stable wall-clock figure (>= 0.1 s), the same number of frames on every in-flight stream
    e2e_steps = max(4 * N_SLOTS, min(4 * args.steps, 960)) // N_SLOTS * N_SLOTS
    run_e2e(2 * N_SLOTS)
    barrier()
    for sl in slots:
        sl["d2h"] = 0
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3          # the host is in this loop: wall clock around fully synchronised ends
    d2h_per_step = sum(sl["d2h"] for sl in slots) // e2e_steps
    barrier()

    # ---- e2e, reference picture kept on the device (SURVEY.md 8f item 4 closed into a loop): per frame only the source goes up;
    # the cost tables, the level streams of the chosen units and the SAO statistics come down; gather -> deblocking -> SAO
    # statistics -> stand-in SAO decision (host) -> SAO offset pass -> border produce the next reference picture in HBM.
    # Every in-flight stream is a real IPPP chain here (frame n+1 searches in the finished frame n).
    from homerhevc_b200.lib import SAO_PARAM_DT
    SAO_LAMBDA, DBK = (float(LAMBDA), LAMBDA / 1.26, LAMBDA / 1.26), (2, 2, 0, 0)
    for sl in slots:
        c = sl["ctx"]
        sl["refs"] = [sl["ref"], hb.Frame(c, w, h)]; sl["rec"] = hb.Frame(c, w, h); sl["which"] = 0; sl["count"] = 0
        sl["levels"] = c.pinned(4 * w * h); sl["prm"] = np.zeros(n_ctus, SAO_PARAM_DT)
        sl["refs"][0].upload_u8(*pinned[0])

    def begin_frame_res(sl, i):
        j = 1 + sl["count"] % N_RESIDENT
        sl["count"] += 1
        sl["pp"].frame_begin_resident(sl["cur"], sl["refs"][sl["which"]], pinned[j], AVG_DIST, sl["tables"])

    def finish_frame_res(sl):
        nlev = sl["pp"].frame_finish_resident(sl["cur"], LAMBDA, sl["tables"], sl["sel"], sl["off"], sl["rec"], sl["refs"][1 - sl["which"]], DBK, SAO_LAMBDA,
                                              sl["levels"], sl["prm"])
        sl["which"] = 1 - sl["which"]
        sl["d2h"] += nlev + sl["tables"].nbytes + n_ctus * 15 * 16
        sl["h2d"] = sl.get("h2d", 0) + frame_bytes + sl["prm"].nbytes + sl["sel"].nbytes + sl["off"].nbytes

    res_ms, res_d2h, res_h2d = None, 0, 0
    try:
        begin_frame, finish_frame = begin_frame_res, finish_frame_res          # run_e2e picks them up by name
        run_e2e(2 * N_SLOTS)
        barrier()
        for sl in slots:
            sl["d2h"] = 0; sl["h2d"] = 0
        l_res0 = sum(sl["ctx"].launch_count() for sl in slots)
        t0 = time.perf_counter()
        run_e2e(e2e_steps)
        torch.cuda.synchronize()
        res_ms = (time.perf_counter() - t0) * 1e3
        res_launches = sum(sl["ctx"].launch_count() for sl in slots) - l_res0
        res_d2h = sum(sl["d2h"] for sl in slots) // e2e_steps
        res_h2d = sum(sl["h2d"] for sl in slots) // e2e_steps
        sao_on = float(np.mean([(sl["prm"]["type"] >= 0).mean() for sl in slots]))
    except hb.HbError as e:
        print(f"[bench] device-resident e2e variant failed: {e}", file=sys.stderr)
        res_ms = None
    barrier()
    # the unfiltered variant for reference: one stream, every table / level / reconstruction of all five passes fetched
    e2e_cur, e2e_ref = hb.Frame(ctx, w, h), hb.Frame(ctx, w, h)
    def step_full(i):
        j = i % N_RESIDENT
        e2e_cur.upload_u8(*pinned[j + 1]); e2e_ref.upload_u8(*pinned[j])
        pp.run(e2e_cur, e2e_ref, AVG_DIST)
        pp.fetch_all(out_pin)
    for i in range(3):
        step_full(i)
    t0 = time.perf_counter()
    for i in range(20):
        step_full(i)
    full_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # ---- per-kernel device times (CUDA events between launches, same stream) for the roofline
    prof = {}
    for rep in range(5):
        for name, t in pp.run_profiled(resident[rep + 1], resident[rep], AVG_DIST):
            prof.setdefault(name, []).append(t)
    prof = {k: statistics.mean(v[1:]) for k, v in prof.items()}

    if world > 1:
        t = torch.tensor([ms, e2e_ms, full_ms, res_ms if res_ms is not None else 1e12], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, full_ms, res_all = t.tolist()
        res_ms = None if res_all >= 1e12 else res_all
    if rank == 0:
        peaks, peak_src = measured_peaks()
        abytes = algorithmic_bytes(pp, w, h)
        top = max(prof, key=prof.get)
        achieved = abytes[top] / (prof[top] * 1e-3) / 1e9
        step_bytes = sum(abytes.values())
        total_prof = sum(prof.values())
        # instruction-issue roofline (the governing one): warp instructions per frame counted by ncu
        # (profiles/inst_r01_v26.json, 1080p; scaled by the pixel count for the other sizes) against SMs x 4 schedulers x clock
        issue, traffic = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "inst_r01_v26.json")) as f:
                prof_counts = json.load(f)
            scale = (w * h) / (1920 * 1080)
            inst_per_frame = prof_counts["total_warp_inst"] * scale
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            peak_issue = 148 * 3.8 * sm_hz        # measured: 3.80 warp-inst/clk/SM when both integer pipes are fed (tools/int_peak.cu, profiles/int_peak_r01.txt)
            fps_gpu = N_SLOTS * args.steps / (ms * 1e-3)
            issue = {"bound": "warp-instruction issue (int-op roofline)", "achieved": inst_per_frame * fps_gpu, "peak": peak_issue, "unit": "warp-inst/s",
                     "frac": inst_per_frame * fps_gpu / peak_issue, "warp_inst_per_frame": inst_per_frame,
                     "source": "ncu smsp__inst_executed.sum per kernel, profiles/inst_r01_v26.json" + ("" if scale == 1 else " (scaled by pixel count)")}
            if scale == 1:
                want = {"me": "k_me<", "mc": "k_mc", "tq": "k_tq<"}[top[:2]] + (top[2:] + ">" if top[:2] == "me" else "")
                for kk in prof_counts["kernels"]:
                    if kk["kernel"].startswith(want):
                        traffic = kk["dram_read_bytes"] + kk["dram_write_bytes"]
                        break
        except Exception:
            pass
        line = {
            "metric": "ME+TQ frames/s", "value": world * N_SLOTS * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int16 residual/levels, int32 accumulate",
            "data": "synthetic",
            "config": {"workload": workload_name(args, w, h), "frames_per_step_per_gpu": N_SLOTS, "streams_in_flight_per_gpu": N_SLOTS, "qp": QP, "avg_dist": AVG_DIST,
                       "parallelism": f"gop-per-gpu x{world}" if world > 1 else "single gpu",
                       "l2": f"inputs rotate over {N_RESIDENT} resident frame pairs; a step touches ~{step_bytes / 2**20:.0f} MiB, "
                             f"{N_RESIDENT} steps > {L2_BYTES / 2**20:.0f} MiB L2 before any input is reused",
                       "cuda_graph": True},
            "e2e": {"value": world * e2e_steps / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 2 * frame_bytes,
                    "d2h_bytes_per_step": int(d2h_per_step), "steps": e2e_steps, "streams_in_flight": N_SLOTS, "host_threads": E2E_THREADS,
                    "flow": "upload cur+ref -> pre-pass -> fetch cost tables (compact ME records + one 12-byte cost record per coding unit and pass) -> host depth choice per CTU -> gather + fetch recon and coded levels of that choice",
                    "fetch_everything_variant": {"value": world * 20 / (full_ms * 1e-3), "unit": "frames/s", "d2h_bytes_per_step": out_bytes},
                    "device_resident_reference_variant": None if res_ms is None else {
                        "value": world * e2e_steps / (res_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(res_h2d), "d2h_bytes_per_step": int(res_d2h),
                        "gpu_launches": int(res_launches), "ctus_with_sao": round(sao_on, 3),
                        "flow": "upload cur only -> pre-pass against the finished previous frame in HBM -> fetch cost tables -> host choice per CTU -> gather into a frame + "
                                "deblocking (strengths from the plan's tables) + fetch coded levels -> SAO statistics + per-type offsets/distortions on the device -> fetch those -> host SAO type choice (stand-in) -> SAO offset pass + border"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms": prof[top], "kernel_share_of_step": prof[top] / total_prof,
                         "step_algorithmic_bytes": step_bytes,
                         "step_frac": N_SLOTS * step_bytes / (ms / args.steps * 1e-3) / 1e9 / peaks["hbm_gbs"]},
            "issue_roofline": issue,
            "kernels_ms": {k: round(v, 5) for k, v in prof.items()},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from _oracle import have_ref, ref_prepass
                if have_ref():
                    cores = len(os.sched_getaffinity(0))
                    ref_prepass(host[1], host[0], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False)     # warm
                    n_cpu, secs = 0, 0.0
                    while secs < 8.0 and n_cpu < 64:
                        s, _ = ref_prepass(host[n_cpu % N_RESIDENT + 1], host[n_cpu % N_RESIDENT], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False)
                        secs += s; n_cpu += 1
                    line["cpu_baseline"] = {"value": n_cpu / secs, "unit": "frames/s", "cores": cores, "kind": "reference",
                                            "sample": f"{n_cpu} frames of the same workload through the reference's own functions (oracle/_ref), {cores} threads"}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
            except Exception as e:  # the baseline is a reported figure; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
