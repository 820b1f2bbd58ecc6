#!/usr/bin/env python
"""bench.py -- ME+TQ frames/s of the frame-level pre-pass (BASELINE.json metric) on N B200s.

A step = one pass of the hot path over one frame pair per in-flight stream: motion search of every PU 64/32/16/8 (parent -> child
chain, integer walk + half + quarter pel), motion compensation luma + chroma at each size, and the inter T/Q chain over TU
32/32/16/8/4 (+ chroma) with reconstruction.  Workload: synthetic YUV 4:2:0 of SURVEY.md 8(d), IPPP quarter-pel, fixed QP 32.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 720p|1080p|2160p] [--impl ours|reference] [--mode ...]

One JSON line (rank 0).  What it carries:
  value        frames/s of the pre-pass with inputs resident in HBM, independent GOP streams per GPU (weak scaling over N);
  e2e          the same through the C ABI with HOST buffers, reference picture kept on the device (source up, cost tables / level
               streams / SAO candidates down, the finished picture becomes the next reference in HBM); the flow that also moves the
               reference picture over the link is `e2e.host_reference_variant`;
  roofline     the dominant kernel's ALGORITHMIC integer work (pixel-abs-diffs of the reference's own search pattern +
               interpolation / transform MACs) / its device time, against the measured integer-op ceiling (tools/int_peak.cu); the HBM
               view of the same launch under `roofline.hbm`; executed-instruction issue utilisation under `issue_roofline` (only from an
               ncu profile taken on exactly these kernel sources);
  bands_2160p  BASELINE.json configs[3]: ONE 3840x2160 frame split into CTU-row bands over the N GPUs, reference halos exchanged
               over NCCL on the device timeline (strong scaling), with the band tables checked against a whole-frame run in-process;
  extra_workloads (N = 1)  value / e2e of the other two picture sizes;
  cpu_baseline (N = 1)  the reference's own functions on the host cores, and its console encoder homer_app in both thread settings.
`--impl reference` times the reference's CPU implementation of the same path (oracle/_ref, compiled from /root/reference); rank 0 only.
`--mode encode` times the reference's whole encoder with the batched API in its loop (oracle/ref_hooks.c) next to homer_app.
"""
import argparse
import hashlib
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
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
N_SM = 148
# measured on a B200 (tools/int_peak.cu, profiles/int_peak_r01.txt): warp instructions per clock and SM
PEAK_BOTH_PIPES, PEAK_ONE_PIPE = 3.80, 1.96


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def kernel_source_sha():
    """identity of the kernel sources: an ncu profile is only quoted when it was taken on exactly these"""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "homerhevc_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode()); h.update(f.read())
    return h.hexdigest()[:16]


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


def workload_name(w, h):
    return f"{w}x{h} IPPP quarter-pel ME (PU 64/32/16/8) + MC + inter T/Q (TU 32/32/16/8/4), fixed QP {QP}, synthetic YUV420"


def bench_config(args, w, h, world, streams):
    """the SAME dictionary in both arms (ours / --impl reference): the workload, not the implementation"""
    return {"workload": workload_name(w, h), "qp": QP, "avg_dist": AVG_DIST,
            "parallelism": f"gop-per-gpu x{world}" if world > 1 else "single gpu",
            "streams_in_flight_per_gpu": streams,
            "l2": f"inputs rotate over {N_RESIDENT} resident frame pairs (a 1080p step touches ~114 MiB per stream): no step finds its inputs in the {L2_BYTES // 2**20} MiB L2"}


# ---------------------------------------------------------------------------------------------------------------- algorithmic work
def algorithmic_bytes(pp, w, h):
    """per-launch algorithmic HBM bytes of every pre-pass kernel: each input plane read once,
    each output written once, for the PUs / TUs that launch covers."""
    out = {}
    luma = w * h
    out["sp"] = luma + 15 * luma                                  # reference luma in, fifteen quarter-pel planes out
    for d in range(4):
        s = 64 >> d
        n_pu = (w // s) * (h // s)
        # cur + ref luma (u8) in, the 16 sub-pel candidate blocks of every PU from the planes, one result record and the luma prediction out
        out[f"me{s}"] = 2 * luma + 16 * n_pu * s * s + 24 * n_pu + n_pu * s * s
        out[f"mc{s}"] = int(0.5 * n_pu * s * s) * 2 + 8 * n_pu     # chroma: ref in, pred out, mv in
    # the single-launch search reads cur + ref once for all four PU sizes
    out["me"] = 2 * luma + sum(out[f"me{64 >> d}"] - 2 * luma for d in range(4))
    for p in range(5):
        for c in range(3):
            t = pp.tu_size(p, c)
            if not t:
                continue
            n = pp.num_tus(p, c)
            out[f"tq{p}{'yuv'[c]}{t}"] = n * t * t * (1 + 1 + 1 + 2) + n * (8 + 16)   # cur, pred in; recon, levels out; job, result
    return out


def algorithmic_ops(pp, w, h):
    """per-launch algorithmic INTEGER work, independent of how a kernel is written:
      search   pixel-abs-diffs of the reference's own pattern = (integer probes the walk makes, counted by the kernel itself, + 18
               sub-pel probes) x N^2 per PU, plus the interpolation MACs of the reference's per-PU plane builders: 13 passes x 8 taps
               over (N+1) x (N+8) samples (hmr_motion_inter.c:395-562) -- charged to the search launches as SURVEY.md 8(d) defines the
               work, although this implementation builds the planes once per picture (`sp`: 15 x 8 MAC per sample);
      T/Q      forward 2-D transform as matrix products 2N MAC per sample, the same again for the inverse of the units that keep
               levels, + 2 multiplies per sample for quantisation / dequantisation;
      chroma MC  two 4-tap passes per sample, both planes.
    -> {kernel: {"pad": .., "mac": .., "packing": ..}}; packing = how many of these operations one lane instruction can carry at best
    (4 for 8-bit samples: VABSDIFF4 / dp4a; 2 for the 16-bit operands of the transforms: dp2a) and on how many of the two integer pipes."""
    out = {"sp": {"pad": 0, "mac": 15 * 8 * w * h, "lanes_per_inst": 4, "pipes": PEAK_ONE_PIPE}}     # 3 H + 12 V passes x 8 taps per sample
    for d in range(4):
        s = 64 >> d
        me = pp.fetch_me(d)
        ok = me["sad"] != 0xFFFFFFFF
        n_pu = int(ok.sum())
        probes = int(me["n_probes"][ok].astype(np.int64).sum())
        out[f"me{s}"] = {"pad": (probes + 18 * n_pu) * s * s, "mac": n_pu * 13 * 8 * (s + 1) * (s + 8), "lanes_per_inst": 4, "pipes": PEAK_BOTH_PIPES}
        out[f"mc{s}"] = {"pad": 0, "mac": n_pu * 2 * (s // 2) * (s // 2) * 8, "lanes_per_inst": 4, "pipes": PEAK_ONE_PIPE}
    # the single-launch search (a CTA per CTU, all four PU sizes): the work of the four per-size launches
    out["me"] = {"pad": sum(out[f"me{64 >> d}"]["pad"] for d in range(4)), "mac": sum(out[f"me{64 >> d}"]["mac"] for d in range(4)),
                 "lanes_per_inst": 4, "pipes": PEAK_BOTH_PIPES}
    for p in range(5):
        for c in range(3):
            t = pp.tu_size(p, c)
            if not t:
                continue
            res = pp.fetch_tu(p, c)
            n, coded = len(res), int((res["sum"] > 0).sum() + (res["zeroed"] != 0).sum())
            out[f"tq{p}{'yuv'[c]}{t}"] = {"pad": 0, "mac": (n + coded) * 2 * t * t * t + 2 * n * t * t, "lanes_per_inst": 2, "pipes": PEAK_ONE_PIPE}
    return out


def int_peak_ops(entry, sm_hz):
    """operations per second the CUDA cores can retire for this kind of work: SMs x warp-inst/clk x 32 lanes x packed ops per lane"""
    return N_SM * entry["pipes"] * 32 * entry["lanes_per_inst"] * sm_hz


# ---------------------------------------------------------------------------------------------------------------- reference arm
def homer_app_fps(w, h, n_frames, wpp, engines, seed=5):
    """the reference's own console encoder (oracle/_ref/homer_app, compiled unmodified) on a synthetic clip: its own fps print"""
    from homerhevc_b200 import synth
    exe = os.path.join(ROOT, "oracle", "_ref", "homer_app")
    if not os.path.exists(exe):
        return None
    clip = synth.make_clip(w, h, n_frames, seed=seed)
    with tempfile.TemporaryDirectory() as td:
        src, dst = os.path.join(td, "in.yuv"), os.path.join(td, "out.265")
        with open(src, "wb") as f:
            for fr in clip:
                for p in fr:
                    f.write(p.tobytes())
        cmd = [exe, "-i", src, "-o", dst, "-widthxheight", f"{w}x{h}", "-gop_size", "1", "-b_frames", "0", "-bitrate_mode", "0", "-qp", str(QP),
               "-n_frames", str(n_frames), "-n_wpp_threads", str(wpp), "-n_enc_engines", str(engines)]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            tail = out.stdout.replace("\r", "\n").strip().splitlines()[-1]
            fps = float(tail.split(":")[-1].split("fps")[0])
            return {"value": fps, "unit": "frames/s", "frames": n_frames, "n_wpp_threads": wpp, "n_enc_engines": engines,
                    "stream_bytes": os.path.getsize(dst) if os.path.exists(dst) else None}
        except Exception as e:
            return {"value": None, "what": f"failed: {e}"}


def homer_app_baseline(w, h, cores):
    """parity setting (1 engine, WPP off) and throughput setting (WPP threads = CTU rows or cores, 3 engines)"""
    nf = 12 if (w, h) == WORKLOADS["720p"] else (8 if (w, h) == WORKLOADS["1080p"] else 4)
    rows = (h + 63) // 64
    return {"nproc": cores, "cmd": "oracle/_ref/homer_app -gop_size 1 -b_frames 0 -bitrate_mode 0 -qp 32 (fixed QP, IPPP, quarter-pel, all stages incl. CABAC)",
            "parity": homer_app_fps(w, h, nf, 0, 1), "throughput": homer_app_fps(w, h, 3 * nf, min(rows, max(cores, 1)), 3)}


def run_reference(args, w, h, rank, world):
    """the reference's own CPU functions over the same pre-pass, all host threads; a step = `sample` frames; seconds are the C
    driver's own clock around its worker threads (output arrays are allocated once, outside)"""
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
            s, _ = ref_prepass(frames[j + 1], frames[j], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False, reuse_outputs=True)
            t += s
        return t
    for i in range(max(args.warmup, 1)):
        step(i)
    secs = sum(step(i) for i in range(args.steps))
    fps = args.steps * sample / secs
    line = {"impl": "reference", "metric": "ME+TQ frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 samples, int16 residual/levels, int32 accumulate", "data": "synthetic",
            "config": bench_config(args, w, h, world, max(1, min(args.value_streams, args.streams))),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": f"{sample} frame(s) per step, the reference's own functions (oracle/_ref: hmr_motion_estimation, hmr_motion_compensation_*, "
                                       f"predict, encode_inter_cu*), {cores} threads, CTUs dealt round-robin, timed inside the C driver"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["homer_app"] = homer_app_baseline(w, h, cores)
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------- side modes
def run_intra(args, w, h, rank, world, local, hb, synth, quiet=False, max_steps=40):
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
    steps = min(args.steps, max_steps)
    t0 = time.perf_counter()
    for i in range(steps):
        sads = ctx.intra_presearch(dev[i % 4], rec, adi_pin[i % 4], sad_pin)
    secs = time.perf_counter() - t0
    jobs, adi, off = prep[(steps - 1) % 4]
    line = {"metric": "intra 35-mode pre-search frames/s", "value": steps / secs, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": max(3, args.warmup),
            "ms_per_step": secs / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int32 accumulate",
            "data": "synthetic", "config": {"workload": workload_name(w, h) + ", every 32/16/8/4 luma block, reference samples from the original picture",
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
    for d in dev:
        d.close()
    ctx.close()
    if not quiet:
        print(json.dumps(line))
    return line


def run_intra_recon(args, w, h, rank, world, local, torch, dist, hb, synth, barrier, quiet=False, n_streams=None, max_steps=24):
    """BASELINE.json configs[4] / [0]: reconstruction of I pictures on the device, wavefront-batched (hb_intra_reconstruct), from the decisions the
    reference's own encoder made for the picture (one reference encode on the CPU, outside the timed region); every GPU runs its own independent
    intra GOPs (streams), one host thread per stream: per picture the source goes up, the units' levels / sums / distortions come down and the
    reconstruction stays on the device.  Weak scaling over N.  Next to it: the oracle's restatement of the same reconstruction on one host core
    and the reference's whole all-intra encode (decisions + entropy coding included)."""
    from _intra import capture_intra_picture, intra_tus, oracle_intra_recon
    from _oracle import have_ref
    if not have_ref():
        if rank == 0 and not quiet:
            print(json.dumps({"metric": "intra reconstruction frames/s", "unavailable": "oracle/_ref (compiled reference) is not in this tree: no decisions to rebuild"}))
        return None
    seed = synth.SEEDS.get((w, h), 7)
    clip = synth.make_clip(w, h, 1, seed=seed)
    cap = capture_intra_picture(w, h, QP, 1, seed)      # own process, zero-initialised build of the reference (its as-is build depends on stack garbage)
    t_ref_encode = cap["seconds"]
    tus = intra_tus(cap, w, h)
    S = n_streams or max(1, min(args.streams, 8))
    slots = []
    for k in range(S):
        c = hb.Context(local)
        slots.append({"ctx": c, "f": [hb.Frame(c, w, h) for _ in range(3)], "pin": [c.pinned(p.nbytes).view(np.uint8).reshape(p.shape) for p in clip[0]]})
        for dst, src in zip(slots[-1]["pin"], clip[0]):
            dst[:] = src
    steps = max(3, min(args.steps, max_steps))
    state = {"levels": 0, "ok": True}

    def work(sl, n):
        try:
            for _ in range(n):
                sl["f"][0].upload_u8(*sl["pin"])
                _, _, lv = sl["ctx"].intra_reconstruct(sl["f"][0], sl["f"][1], sl["f"][2], tus, 1, 1, 1.0)
                state["levels"] = lv
        except Exception as e:          # a dead worker must not turn into a fast, wrong number
            state["error"] = e

    def run(n):
        th = [threading.Thread(target=work, args=(sl, n)) for sl in slots]
        for t in th: t.start()
        for t in th: t.join()
        if state.get("error"):
            raise state["error"]
    run(2)
    got = slots[0]["f"][2].download()
    identical = all(np.array_equal(got[c], cap["recon"][c]) for c in range(3))
    barrier()
    t0 = time.perf_counter()
    run(steps)
    secs = time.perf_counter() - t0
    barrier()
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = t.item()
    launches = sum(int(sl["ctx"].launch_count()) for sl in slots)
    line = None
    if rank == 0:
        fps = world * S * steps / secs
        line = {"metric": "intra reconstruction frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": 2, "ms_per_step": 1e3 * secs / steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int16 residual/levels", "data": "synthetic",
                "config": {"workload": workload_name(w, h).split(" IPPP")[0] + f" I picture, fixed QP {QP}: {len(tus)} transform units of the reference encoder's own decisions in {state['levels']} dependency levels",
                           "parallelism": f"{S} independent intra GOPs per GPU, one host thread each" + (f", x{world} GPUs" if world > 1 else ""),
                           "timing": "wall clock around the blocking calls, source upload and level download inside"},
                "identical_to_reference_reconstruction": bool(identical),
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": int(w * h * 3 // 2 + tus.nbytes), "d2h_bytes_per_step": int(2 * (tus[:, 3].astype(np.int64) ** 2).sum() + 16 * len(tus))},
                "gpu_launches": launches}
        if not args.no_cpu_baseline:
            t0 = time.perf_counter()
            oracle_intra_recon(clip[0], w, h, tus)
            t_port = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": 1.0 / t_port, "unit": "frames/s", "cores": 1, "kind": "port", "sample": "1 picture through the oracle's restatement of the same reconstruction (oracle/hb_oracle.c: orc_intra_recon_tus)",
                                    "reference_whole_intra_encode_fps": 1.0 / t_ref_encode,
                                    "reference_whole_intra_encode": "the reference's encoder on that picture, decisions and entropy coding included, 1 engine, WPP off (where the decisions come from)"}
        if not quiet:
            print(json.dumps(line))
    for sl in slots:
        for f in sl["f"]:
            f.close()
        sl["ctx"].close()
    return line


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


def run_encode(args, w, h, rank):
    """the reference's WHOLE encoder (host mode decision, CABAC, in-loop filters: unmodified) with the batched GPU API in its loop
    (oracle/ref_hooks.c: hb_enc_me / hb_enc_predict / hb_enc_tq per coding unit, per-call table for the rest) in lock step, its
    output compared byte for byte with the unmodified encoder's, next to homer_app in both thread settings.  Rank 0 only."""
    if rank != 0:
        return
    import subprocess
    from _oracle import have_ref
    if not have_ref():
        print(json.dumps({"metric": "encoder frames/s", "unavailable": "oracle/_ref (compiled reference) is not in this tree"}))
        return
    nf = max(3, min(args.steps, 10))
    # both arms on oracle/_ref/zinit (the unmodified sources compiled with -ftrivial-auto-var-init=zero), in a process of their own: the as-is
    # build's stream depends on what earlier calls left on the stack (its SSE4.2 intra predictors read automatic variables they never wrote;
    # DESIGN.md section 7, tests/test_gpu_encode_hooks.py), which no replacement can reproduce
    zinit = os.path.join(ROOT, "oracle", "_ref", "zinit")
    ref_dir = zinit if os.path.exists(os.path.join(zinit, "librefdrv.so")) else os.path.join(ROOT, "oracle", "_ref")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "encode_check.py"), f"{w}x{h}x{nf}", "-1", "hooks", "1"], capture_output=True, text=True,
                         timeout=3000, cwd=ROOT, env=dict(os.environ, HB_REF_DIR=ref_dir))
    if out.returncode != 0:
        print(json.dumps({"metric": "encoder frames/s", "unavailable": "tools/encode_check.py failed: " + out.stderr[-300:]}))
        return
    r = json.loads(out.stdout.strip().splitlines()[-1])
    cnt, t_gpu, t_cpu = r["hook_calls"], r["seconds_replaced"], r["seconds_reference"]
    identical, stream_bytes = bool(r["identical"]), int(r["bytes"])
    cores = len(os.sched_getaffinity(0))
    line = {"metric": "encoder frames/s (whole encode, host decisions + CABAC on the CPU, ME / MC / inter T/Q through the batched GPU API per coding unit)",
            "value": nf / t_gpu, "unit": "frames/s", "n_gpus": 1, "steps": nf, "warmup": 0, "ms_per_step": 1e3 * t_gpu / nf, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int16 residual/levels", "data": "synthetic",
            "config": {"workload": f"{w}x{h} IPPP quarter-pel, fixed QP {QP}, {nf} frames, 1 engine, WPP off, lock step", "hook_calls": cnt},
            "identical_stream": identical, "stream_bytes": stream_bytes, "reference_build": os.path.relpath(ref_dir, ROOT),
            "e2e": {"value": nf / t_gpu, "unit": "frames/s", "h2d_bytes_per_step": int(2 * w * h * 3), "d2h_bytes_per_step": 0,
                    "note": "one blocking GPU round trip per hmr_motion_estimation / motion compensation / transform unit call: launch latency, not throughput, sets this number"},
            "gpu_launches": int(cnt["me"] + 3 * cnt["mc"] + 2 * cnt["tq"]),
            "cpu_baseline": {"value": nf / t_cpu, "unit": "frames/s", "cores": 1, "kind": "reference",
                             "sample": f"the same {nf} frames through the unmodified reference library in lock step, 1 engine, WPP off"},
            "homer_app": homer_app_baseline(w, h, cores)}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------- bands (configs[3])
def measure_bands(args, rank, world, local, torch, dist, hb, synth, barrier, steps):
    """BASELINE.json configs[3]: 3840x2160 frames split into CTU-row bands over the GPUs, the reference rows a band needs from the
    other bands (68 luma / 36 chroma rows per side) exchanged over NCCL on the device timeline before every frame's search.
    Strong scaling: the job (S streams x K frames of 2160p) is fixed, every GPU does its band of every frame.  Also checks, in this
    process, that the band tables equal the rows of a whole-frame run on the exchanged reference."""
    from homerhevc_b200 import bands
    w, h = WORKLOADS["2160p"]
    ctu_rows = (h + 63) // 64
    row0, nrows = bands.band_ctu_rows(ctu_rows, world, rank)
    dev = torch.device("cuda", local)
    tex = synth.make_texture(w, h)
    n_pairs = 3
    host = [synth.make_frame(tex, w, h, n) for n in range(n_pairs + 1)]
    exchange = os.environ.get("HB_BANDS_EXCHANGE", "peer")       # peer: CUDA IPC pulls out of the neighbours' HBM; nccl: staged send / recv
    out = {}
    # every rank holds only ITS rows of a reference picture (it reconstructed them); the rest arrives over NVLink
    own_frames = []
    for p in host:
        y0, y1 = bands.band_sample_rows(h, ctu_rows, world, rank)
        own = [np.zeros_like(p[0]), np.zeros_like(p[1]), np.zeros_like(p[2])]
        own[0][y0:y1] = p[0][y0:y1]; own[1][y0 // 2:y1 // 2] = p[1][y0 // 2:y1 // 2]; own[2][y0 // 2:y1 // 2] = p[2][y0 // 2:y1 // 2]
        own_frames.append(own)
    for S in (1, 8, 32):          # frame streams in flight: a band's launches are small, several streams fill the SMs they leave idle
        slots = []
        for k in range(S):
            c = hb.Context(local)
            frames = [hb.Frame(c, w, h) for _ in range(n_pairs + 1)]
            for f, own in zip(frames, own_frames):
                f.upload_u8(*own)
            pp = hb.Prepass(c, w, h, qp=QP, use_graph=1, band=(row0, nrows))
            ex = None
            if world > 1 and exchange == "peer":
                c.sync()
                ex = bands.PeerHaloPuller(hb, dist, c, frames, w, h, world, rank)
                for j in range(len(frames)):
                    ex.mark_ready(j)
            elif world > 1:
                ex = bands.FrameHaloExchanger(torch, dist, c, w, h, world, rank, dev)
            # the current picture is read inside the band only, so the banded upload above is all a rank needs of it
            slots.append({"ctx": c, "frames": frames, "pp": pp, "ex": ex})

        def step(i):
            for sl in slots:
                j = i % n_pairs
                if sl["ex"] and exchange == "peer":
                    sl["ex"].pull(j)                          # wait for the owners' events, one copy kernel out of their HBM, border: all on the stream
                elif sl["ex"]:
                    sl["ex"].exchange(sl["frames"][j])        # queued on the context's stream: export -> NCCL send/recv -> import -> border
                sl["pp"].run(sl["frames"][j + 1], sl["frames"][j], AVG_DIST)

        def sync_all():
            for sl in slots:
                sl["ctx"].sync()
            torch.cuda.synchronize()

        for i in range(max(3, n_pairs)):
            step(i)
        sync_all()
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            step(i)
        sync_all()
        ms = (time.perf_counter() - t0) * 1e3
        barrier()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        out[f"streams_{S}"] = {"value": S * steps / (ms * 1e-3), "unit": "frames/s", "ms_per_frame": ms / (S * steps), "frames": S * steps}
        if S == 1:
            # parity inside the run: this rank's band tables against the same rows of a whole-frame pre-pass on a complete reference
            sl = slots[0]
            whole_ref, whole_cur = hb.Frame(sl["ctx"], w, h), hb.Frame(sl["ctx"], w, h)
            whole_ref.upload_u8(*host[0]); whole_cur.upload_u8(*host[1])
            if sl["ex"] and exchange == "peer":
                sl["ex"].pull(0)
            elif sl["ex"]:
                sl["ex"].exchange(sl["frames"][0])
            sl["pp"].run(sl["frames"][1], sl["frames"][0], AVG_DIST)
            sl["ctx"].sync()
            band_me = [sl["pp"].fetch_me(d) for d in range(4)]
            band_tu = {(p, c): sl["pp"].fetch_tu(p, c) for p in range(5) for c in range(3) if sl["pp"].tu_size(p, c)}
            band_xy = {k: sl["pp"].tu_xy(*k) for k in band_tu}
            full = hb.Prepass(sl["ctx"], w, h, qp=QP, use_graph=0)
            full.run(whole_cur, whole_ref, AVG_DIST)
            sl["ctx"].sync()
            mism = 0
            for d in range(4):
                s = 64 >> d
                fm = full.fetch_me(d)
                gw = ((w + 63) // 64) * (64 // s)
                rows_lo, rows_hi = row0 * (64 // s), (row0 + nrows) * (64 // s)
                a = band_me[d].reshape(-1, gw)[rows_lo:rows_hi]; b = fm.reshape(-1, gw)[rows_lo:rows_hi]
                mism += int((a.tobytes() != b.tobytes()))
            for k, res in band_tu.items():
                if not len(res):
                    continue
                fxy = np.asarray(full.tu_xy(*k), dtype=np.int64).reshape(-1, 2); fres = full.fetch_tu(*k)
                bxy = np.asarray(band_xy[k], dtype=np.int64).reshape(-1, 2)
                fkey, bkey = fxy[:, 1] * 65536 + fxy[:, 0], bxy[:, 1] * 65536 + bxy[:, 0]
                order = np.argsort(fkey)
                idx = order[np.searchsorted(fkey[order], bkey)]
                mism += int(not np.array_equal(fkey[idx], bkey) or res.tobytes() != fres[idx].tobytes())
            if world > 1:
                t = torch.tensor([mism], dtype=torch.int64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                mism = int(t.item())
            out["mismatches"] = mism
            out["parity"] = "every rank: ME tables of its band rows (4 depths) and TU records (13 passes x planes) of a band run on the exchanged reference == the same rows of a whole-frame run on the complete reference"
            out["halo_bytes_sent_per_frame_rank0"] = sl["ex"].bytes_per_exchange if sl["ex"] else 0
            full.close(); whole_ref.close(); whole_cur.close()
        for sl in slots:
            sl["ctx"].sync()
            if sl["ex"] and exchange == "peer":
                sl["ex"].close()                  # views of the neighbours' pictures go before anybody frees a picture
        barrier()
        for sl in slots:
            sl["pp"].close()
            for f in sl["frames"]:
                f.close()
            sl["ctx"].close()
    out.update({"n_gpus": world, "scaling": "strong", "workload": workload_name(w, h),
                "parallelism": f"ctu-row bands x{world}: rank r owns CTU rows [{row0}, {row0 + nrows}) of {ctu_rows}; " + ("halo rows pulled out of the neighbours' HBM over NVLink (CUDA IPC views, one copy kernel per picture on the library's stream, inter-process events; no host wait)" if exchange == "peer" else "halo exchange by grouped NCCL send/recv ordered on the library's stream (no host wait)"),
                "halo_rows": {"luma": bands.HALO_LUMA, "chroma": bands.HALO_CHROMA},
                "timing": "wall clock between device-synchronised barriers, MAX over ranks"})
    return out


# ---------------------------------------------------------------------------------------------------------------- the main measurement
def measure_workload(args, w, h, rank, world, local, torch, dist, hb, synth, barrier, steps, warmup, n_slots, full, n_value=None):
    """value (resident inputs, device-timed), e2e (host buffers, device-resident reference picture) and -- full only -- the other
    e2e variants and the per-kernel profile of one picture size"""
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
    ctx.sync()

    # n_slots independent streams of frames (GOPs, BASELINE.json configs[4]) are in flight per GPU: each has its own context
    # (CUDA stream), pre-pass plan and output buffers, so the search chain of one frame overlaps the T/Q tail of another
    LAMBDA = 60
    slots = []
    for k in range(n_slots):
        c = hb.Context(local)
        spp = hb.Prepass(c, w, h, qp=QP, use_graph=1, compact_tables=2)    # compact ME records + one cost record per coding unit on the wire (e2e); `value` never fetches
        n_ctus = spp.num_ctus()
        slots.append({"ctx": c, "cur": hb.Frame(c, w, h), "ref": hb.Frame(c, w, h), "pp": spp, "tables": c.pinned(spp.tables_bytes()),
                      "out": c.pinned(frame_bytes + 4 * w * h), "sel": np.zeros(n_ctus, np.uint8), "off": np.zeros(n_ctus + 1, np.int32), "d2h": 0})

    # `value` keeps n_value of the streams in flight (default 4: the device is full from four frames on; with more, the kernels of more frames
    # evict each other's instructions and lines -- measured 7 196 / 7 021 / 6 909 frames/s at 4 / 8 / 16 streams); the e2e flows use all of them,
    # they need the depth to cover the host's share of a frame
    n_value = max(1, min(n_value or n_slots, n_slots))
    vslots = slots[:n_value]

    def step_resident(i):
        """one step = one frame on every in-flight stream, inputs resident in HBM"""
        for k, sl in enumerate(vslots):
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
    for i in range(warmup):
        step_resident(i)
    sync_all()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(sl["ctx"].launch_count() for sl in slots)
    ctx.timer_begin()
    for sl in vslots:
        sl["ctx"].wait(ctx)
    for i in range(steps):
        step_resident(i)
    for sl in vslots:
        ctx.wait(sl["ctx"])
    ms = ctx.timer_end()
    launches = sum(sl["ctx"].launch_count() for sl in slots) - l0
    barrier()
    clocks = sampler.stop()
    # ---- one dependent IPPP chain (a single stream, each frame queued behind the previous one): what one encoder instance sees
    one = slots[0]
    one["ctx"].timer_begin()
    n_chain = max(8, min(steps, 64))
    for i in range(n_chain):
        one["pp"].run(resident[i % N_RESIDENT + 1], resident[i % N_RESIDENT], AVG_DIST)
    chain_ms = one["ctx"].timer_end()
    barrier()

    # ---- e2e: the call sequence a host encoder makes, with HOST buffers, per frame.  One host thread per TWO in-flight streams
    # (the reference runs one pthread per encoder engine, hmr_encoder_lib.c:1647): a thread queues frame n+1 on its second stream
    # (the begin call returns at once) before it blocks in the finish call of frame n.  The C calls release the GIL.
    E2E_THREADS = int(os.environ.get("HB_E2E_THREADS", max(1, n_slots // 2)))
    flow = {}

    def run_e2e(n):
        def worker(k):
            mine = [slots[k], slots[k + E2E_THREADS]] if k + E2E_THREADS < n_slots else [slots[k]]
            pending = None
            for c, i in enumerate(range(k, n, E2E_THREADS)):
                sl = mine[c % len(mine)]
                if pending is sl:                       # single-stream fallback: finish before reusing the stream
                    flow["finish"](pending); pending = None
                flow["begin"](sl, i)
                if pending is not None:
                    flow["finish"](pending)
                pending = sl
            if pending is not None:
                flow["finish"](pending)
        ths = [threading.Thread(target=worker, args=(k,)) for k in range(E2E_THREADS)]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    def timed_e2e():
        """warm run (also calibrates the length), then a window of at least 0.6 s whatever --steps says"""
        run_e2e(2 * n_slots)                               # first frames: lazy allocations, graph captures
        torch.cuda.synchronize()
        n = 8 * n_slots
        for attempt in range(4):                           # lengthen the window until it is at least half a second (every rank the same n)
            barrier()
            for sl in slots:
                sl["d2h"] = 0; sl["h2d"] = 0
            l_0 = sum(sl["ctx"].launch_count() for sl in slots)
            t0 = time.perf_counter()
            run_e2e(n)
            torch.cuda.synchronize()
            e_ms = (time.perf_counter() - t0) * 1e3          # the host is in this loop: wall clock around fully synchronised ends
            barrier()
            worst = e_ms
            if world > 1:
                tt = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MIN)
                worst = tt.item()
            if worst >= 500.0 or attempt == 3:
                break
            n = int(min(40000, math.ceil(n * 650.0 / max(worst, 1.0)))) // n_slots * n_slots + n_slots
        return n, e_ms, sum(sl["d2h"] for sl in slots) // n, sum(sl["h2d"] for sl in slots) // n, sum(sl["ctx"].launch_count() for sl in slots) - l_0

    # (1) headline: the reference picture stays on the device (SURVEY.md 8f item 4 closed into a loop): per frame only the source goes
    # up; the cost tables, the level streams of the chosen units and the SAO candidates come down; gather -> deblocking -> SAO
    # statistics -> stand-in SAO decision (host) -> SAO offset pass -> border produce the next reference picture in HBM.  Every
    # in-flight stream is a real IPPP chain here (frame n+1 searches in the finished frame n).
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

    res = None
    try:
        flow["begin"], flow["finish"] = begin_frame_res, finish_frame_res
        n, e_ms, d2h, h2d, nl = timed_e2e()
        res = {"steps": n, "ms": e_ms, "d2h": d2h, "h2d": h2d, "launches": nl, "sao_on": float(np.mean([(sl["prm"]["type"] >= 0).mean() for sl in slots]))}
    except hb.HbError as e:
        print(f"[bench] device-resident e2e flow failed: {e}", file=sys.stderr)

    # (2) the reference picture crosses the link as well: upload cur + ref -> pre-pass -> cost tables -> host choice -> gather + fetch
    # the reconstruction and coded levels of that choice
    def begin_frame(sl, i):
        j = i % N_RESIDENT
        sl["pp"].frame_begin(sl["cur"], sl["ref"], pinned[j + 1], pinned[j], AVG_DIST, sl["tables"])

    def finish_frame(sl):
        sl["d2h"] += sl["pp"].frame_finish(LAMBDA, sl["tables"], sl["sel"], sl["off"], sl["out"]) + sl["tables"].nbytes
        sl["h2d"] = sl.get("h2d", 0) + 2 * frame_bytes

    hostref = None
    if full:
        flow["begin"], flow["finish"] = begin_frame, finish_frame
        n, e_ms, d2h, h2d, nl = timed_e2e()
        hostref = {"steps": n, "ms": e_ms, "d2h": d2h, "h2d": h2d}

    # ---- per-kernel device times (CUDA events between launches, same stream) and the algorithmic work of each launch
    prof, abytes, aops = {}, None, None
    if full:
        for rep in range(5):
            for name, t in pp.run_profiled(resident[rep + 1], resident[rep], AVG_DIST):
                prof.setdefault(name, []).append(t)
        prof = {k: statistics.mean(v[1:]) for k, v in prof.items()}
        abytes = algorithmic_bytes(pp, w, h)
        aops = algorithmic_ops(pp, w, h)          # from the tables the last profiled run left (probe counts, coded flags)
    t = [ms, chain_ms, res["ms"] if res else 1e12, hostref["ms"] if hostref else 1e12]
    if world > 1:
        tt = torch.tensor(t, dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = tt.tolist()
    ms, chain_ms = t[0], t[1]
    if res:
        res["ms"] = t[2]
    if hostref:
        hostref["ms"] = t[3]
    r = {"w": w, "h": h, "ms": ms, "steps": steps, "n_slots": n_slots, "n_value": n_value, "launches": launches, "clocks": clocks, "chain_fps": n_chain / (chain_ms * 1e-3),
         "resident": res, "hostref": hostref, "prof": prof, "abytes": abytes, "aops": aops, "frame_bytes": frame_bytes, "e2e_threads": E2E_THREADS,
         "host_frames": host}
    # release the device memory of this picture size before the next one is set up
    for sl in slots:
        sl["pp"].close()
        for f in [sl["cur"], sl["rec"]] + sl["refs"]:
            f.close()
        sl["ctx"].close()
    pp.close()
    for f in resident:
        f.close()
    ctx.close()
    return r


def e2e_object(r, world):
    res, hostref = r["resident"], r["hostref"]
    if res is None:
        return {"value": None, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "what": "device-resident flow failed, see stderr"}
    e = {"value": world * res["steps"] / (res["ms"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(res["h2d"]), "d2h_bytes_per_step": int(res["d2h"]),
         "steps": res["steps"], "window_s": round(res["ms"] * 1e-3, 3), "streams_in_flight": r["n_slots"], "host_threads": r["e2e_threads"],
         "gpu_launches": int(res["launches"]), "ctus_with_sao": round(res["sao_on"], 3),
         "flow": "per frame and stream: upload the source (pinned host planes) -> pre-pass against the finished previous picture in HBM -> fetch cost tables (compact ME "
                 "records + one 12-byte record per coding unit and pass) -> host depth choice per CTU -> gather into a frame + deblocking + fetch the coded level streams -> "
                 "SAO statistics + per-type offsets on the device -> fetch those -> host SAO type choice (stand-in) -> SAO offset pass + border = the next reference picture"}
    if hostref:
        e["host_reference_variant"] = {"value": world * hostref["steps"] / (hostref["ms"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(hostref["h2d"]),
                                       "d2h_bytes_per_step": int(hostref["d2h"]), "steps": hostref["steps"],
                                       "flow": "upload cur + ref -> pre-pass -> cost tables -> host choice -> gather + fetch reconstruction and coded levels (the round-1 headline)"}
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip extra_workloads and bands_2160p (kernel work, profiling)")
    ap.add_argument("--streams", type=int, default=16, help="independent GOP streams in flight per GPU (e2e flows)")
    ap.add_argument("--value-streams", type=int, default=4, help="of those, the streams the device-timed `value` keeps in flight")
    ap.add_argument("--mode", default="gops", choices=["gops", "bands", "intra", "intra_recon", "finalise", "encode"],
                    help="gops: the headline line (independent GOP streams per GPU + the 2160p band line); bands: only the 2160p CTU-row-band "
                         "measurement; intra / finalise / encode: side measurements (rank 0)")
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
    if args.mode == "encode":
        run_encode(args, w, h, rank)
        return

    import torch
    import torch.distributed as dist
    import homerhevc_b200 as hb
    from homerhevc_b200 import synth

    if os.environ.get("HB_PIN_CPUS") == "1" and world > 1:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(1, len(cpus) // world)
        os.sched_setaffinity(0, set(cpus[local * per:(local + 1) * per]))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def finish():
        if world > 1:
            dist.destroy_process_group()

    if args.mode == "finalise":
        run_finalise(args, w, h, rank, world, local, hb, synth); finish(); return
    if args.mode == "intra":
        run_intra(args, w, h, rank, world, local, hb, synth); finish(); return
    if args.mode == "intra_recon":
        torch.cuda.set_device(local)
        run_intra_recon(args, w, h, rank, world, local, torch, dist, hb, synth, barrier); finish(); return
    if args.mode == "bands":
        torch.cuda.set_device(local)
        b = measure_bands(args, rank, world, local, torch, dist, hb, synth, barrier, max(8, min(args.steps, 40)))
        if rank == 0:
            print(json.dumps({"metric": "ME+TQ frames/s", "value": b["streams_8"]["value"], "unit": "frames/s", "n_gpus": world, "steps": b["streams_8"]["frames"],
                              "warmup": 3, "ms_per_step": b["streams_8"]["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                              "dtype": "u8 samples, int16 residual/levels, int32 accumulate", "data": "synthetic",
                              "config": {"workload": b["workload"], "parallelism": b["parallelism"]}, "bands_2160p": b}))
        finish(); return

    torch.cuda.set_device(local)
    n_slots = max(1, args.streams)
    n_value = max(1, min(args.value_streams, n_slots))
    r = measure_workload(args, w, h, rank, world, local, torch, dist, hb, synth, barrier, args.steps, args.warmup, n_slots, True, n_value)
    extras = {}
    if world == 1 and not args.no_extras:
        for name, (ew, eh) in WORKLOADS.items():
            if (ew, eh) == (w, h):
                continue
            try:
                x = measure_workload(args, ew, eh, rank, world, local, torch, dist, hb, synth, barrier, max(8, min(args.steps, 40)), 3, n_slots, False, n_value)
                extras[name] = {"value": n_value * x["steps"] / (x["ms"] * 1e-3), "unit": "frames/s", "e2e": e2e_object(x, world)["value"],
                                "one_stream_chain": x["chain_fps"], "steps": x["steps"]}
            except Exception as e:
                extras[name] = {"value": None, "what": f"failed: {e}"}
    if world == 1 and rank == 0 and not args.no_extras:
        # BASELINE.json configs[0] / [4] are intra: the 35-mode pre-search of every 32/16/8/4 luma block of a 1080p picture
        try:
            il = run_intra(args, 1920, 1080, rank, world, local, hb, synth, quiet=True, max_steps=12)
            extras["intra_presearch_1080p"] = {"value": il["value"], "unit": "frames/s", "blocks": il["config"]["blocks"], "modes": 35,
                                               "cpu_reference": (il.get("cpu_baseline") or {}).get("value"), "cpu_cores": (il.get("cpu_baseline") or {}).get("cores"),
                                               "identical_to_reference": (il.get("cpu_baseline") or {}).get("identical_to_gpu")}
        except Exception as e:
            extras["intra_presearch_1080p"] = {"value": None, "what": f"failed: {type(e).__name__}: {e}"}
        # ... and the wavefront-batched reconstruction of an I picture from the reference encoder's decisions (configs[0]: 720p)
        try:
            rl = run_intra_recon(args, 1280, 720, rank, world, local, torch, dist, hb, synth, barrier, quiet=True, max_steps=10)
            if rl:
                extras["intra_reconstruction_720p"] = {"value": rl["value"], "unit": "frames/s", "config": rl["config"], "identical_to_reference_reconstruction": rl["identical_to_reference_reconstruction"],
                                                       "cpu_port_1core": (rl.get("cpu_baseline") or {}).get("value"), "reference_whole_intra_encode_fps": (rl.get("cpu_baseline") or {}).get("reference_whole_intra_encode_fps")}
        except Exception as e:
            extras["intra_reconstruction_720p"] = {"value": None, "what": f"failed: {type(e).__name__}: {e}"}
    bands = None
    if not args.no_extras:
        try:
            bands = measure_bands(args, rank, world, local, torch, dist, hb, synth, barrier, max(8, min(args.steps, 24)))
        except Exception as e:
            bands = {"value": None, "what": f"failed: {type(e).__name__}: {e}"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        prof, abytes, aops, ms, clocks = r["prof"], r["abytes"], r["aops"], r["ms"], r["clocks"]
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
        top = max(prof, key=prof.get)
        total_prof = sum(prof.values())
        ops_top = aops[top]["pad"] + aops[top]["mac"]
        peak_top = int_peak_ops(aops[top], sm_hz)
        achieved = ops_top / (prof[top] * 1e-3)
        # the whole step: time the integer pipes need at best for every launch's algorithmic work / the measured step time per frame
        ideal_s = sum((aops[k]["pad"] + aops[k]["mac"]) / int_peak_ops(aops[k], sm_hz) for k in prof if k in aops)
        frame_s = ms * 1e-3 / (args.steps * n_value)
        step_bytes = sum(abytes[k] for k in prof if k in abytes)
        # executed-instruction view and DRAM traffic: only from an ncu profile taken on exactly these kernel sources
        issue, traffic, sha = None, None, kernel_source_sha()
        try:
            with open(os.path.join(ROOT, "profiles", "inst_r02.json")) as f:
                pc = json.load(f)
            if pc.get("kernel_source_sha") != sha:
                issue = {"stale_profile": True, "profile_sha": pc.get("kernel_source_sha"), "kernel_source_sha": sha}
            elif (w, h) == WORKLOADS["1080p"]:
                inst = pc["total_warp_inst"]
                fps_gpu = n_value * args.steps / (ms * 1e-3)
                issue = {"what": "executed warp instructions (ncu smsp__inst_executed.sum, profiles/inst_r02.json) x frames/s against the measured issue ceiling; utilisation, not a roofline",
                         "warp_inst_per_frame": inst, "achieved": inst * fps_gpu, "peak": N_SM * PEAK_BOTH_PIPES * sm_hz, "unit": "warp-inst/s",
                         "frac": inst * fps_gpu / (N_SM * PEAK_BOTH_PIPES * sm_hz), "kernel_source_sha": sha}
                want = "k_me_ctu" if top == "me" else {"me": "k_me<", "mc": "k_mc", "tq": "k_tq", "sp": "k_subpel"}[top[:2]]
                for kk in pc["kernels"]:
                    if kk["kernel"].startswith(want) and (top[:2] != "me" or top == "me" or kk["kernel"].startswith(f"k_me<{top[2:]},")):
                        traffic = (kk.get("dram_read_bytes") or 0) + (kk.get("dram_write_bytes") or 0)
                        break
        except FileNotFoundError:
            issue = {"stale_profile": True, "what": "no profiles/inst_r02.json for these kernel sources", "kernel_source_sha": sha}
        except Exception as e:
            issue = {"stale_profile": True, "what": f"{type(e).__name__}: {e}"}
        e2e = e2e_object(r, world)
        line = {
            "metric": "ME+TQ frames/s", "value": world * n_value * args.steps / (ms * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 samples, int16 residual/levels, int32 accumulate",
            "data": "synthetic",
            "config": bench_config(args, w, h, world, n_value),
            "gpu_launches": int(r["launches"]),
            "clocks": clocks,
            "kernels_ms": {k: round(v, 5) for k, v in prof.items()},
            "issue_roofline": issue,
            "roofline": {"bound": "int-op", "kernel": top, "achieved": achieved / 1e9, "peak": peak_top / 1e9, "unit": "Gop/s", "frac": achieved / peak_top,
                         "traffic": traffic,
                         "what": "algorithmic integer operations of the launch (pixel-abs-diffs of the reference's own search pattern, probe counts from the kernel's own "
                                 "result table, + interpolation MACs of the reference's per-PU plane builders; SURVEY.md 8d) / its device time (CUDA events around the launch, "
                                 "same stream), against SMs x measured warp-inst/clk (tools/int_peak.cu: both integer pipes 3.80, one 1.96) x 32 lanes x 4 packed 8-bit operations",
                         "algorithmic_ops": {"pad": aops[top]["pad"], "mac": aops[top]["mac"]}, "kernel_ms": prof[top], "kernel_share_of_step": prof[top] / total_prof,
                         "step_frac": ideal_s / frame_s, "peak_source": "profiles/int_peak_r01.txt (measured on a B200)",
                         "hbm": {"achieved": abytes[top] / (prof[top] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": abytes[top] / (prof[top] * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": abytes[top], "peak_source": peak_src,
                                 "step_algorithmic_bytes": step_bytes, "step_frac": step_bytes / frame_s / 1e9 / peaks["hbm_gbs"]}},
            "e2e": e2e,
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                from _oracle import have_ref, ref_prepass
                if have_ref():
                    cores = len(os.sched_getaffinity(0))
                    host = r["host_frames"]
                    ref_prepass(host[1], host[0], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False, reuse_outputs=True)     # warm
                    n_cpu, secs = 0, 0.0
                    while secs < 8.0 and n_cpu < 64:
                        s, _ = ref_prepass(host[n_cpu % N_RESIDENT + 1], host[n_cpu % N_RESIDENT], w, h, QP, AVG_DIST, n_threads=cores, want_pred=False, reuse_outputs=True)
                        secs += s; n_cpu += 1
                    line["cpu_baseline"] = {"value": n_cpu / secs, "unit": "frames/s", "cores": cores, "kind": "reference",
                                            "sample": f"{n_cpu} frames of the same workload through the reference's own functions (oracle/_ref), {cores} threads, timed inside the C driver",
                                            "homer_app": homer_app_baseline(w, h, cores)}
                else:
                    line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
            except Exception as e:  # the baseline is a reported figure; never lose the GPU line over it
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        if extras:
            line["extra_workloads"] = extras
        line["bands_2160p"] = bands
        # short figures at the very end of the line (a truncated tail still carries them)
        line["summary"] = {"n_gpus": world, "value_fps": round(line["value"], 1), "e2e_fps": None if e2e["value"] is None else round(e2e["value"], 1),
                           "e2e_host_ref_fps": round(e2e["host_reference_variant"]["value"], 1) if "host_reference_variant" in e2e else None,
                           "one_stream_chain_fps": round(r["chain_fps"], 1), "roofline_frac": round(line["roofline"]["frac"], 4),
                           "roofline_step_frac": round(line["roofline"]["step_frac"], 4),
                           "bands_2160p_fps_1stream": None if not bands or "streams_1" not in bands else round(bands["streams_1"]["value"], 1),
                           "bands_2160p_fps_8streams": None if not bands or "streams_8" not in bands else round(bands["streams_8"]["value"], 1),
                           "bands_2160p_fps_32streams": None if not bands or "streams_32" not in bands else round(bands["streams_32"]["value"], 1),
                           "bands_mismatches": None if not bands else bands.get("mismatches")}
        print(json.dumps(line))
    finish()


if __name__ == "__main__":
    main()
