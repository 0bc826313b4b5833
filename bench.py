#!/usr/bin/env python
"""bench.py — loop-closure queries/sec on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus 1 --steps 30 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU algorithm on the host cores

Workload = BASELINE.json configs[1]: 640x480 synthetic RGB-D stream, ORB 1000 kp/frame, 49 152-word binary
dictionary, 10 000 signatures.  One "step" = one batch of B independent loop-closure queries (frames); each
frame goes through the WHOLE hot path:
  detect   BGR->gray, depth mask, ORB detect + describe, 3-D lifting      (Memory::createSignature feature block)
  quantise exact 2-NN + NNDR + intra-frame new words against the dictionary (VWDictionary::addNewWords)
  score    TF-IDF over the inverted index of all signatures                (Memory::computeLikelihood)
  verify   top hypothesis: descriptor matching + PnP RANSAC + refinement   (Memory::computeTransform)
The map is built (untimed) from 50 textured places seen 200 times each; the vocabulary is made of the
places' own ORB descriptors, query frames are shifted, noisy re-observations of random places, so the
likelihood arg-max is a true loop closure and its verification succeeds.

`value`   frames/s with images + depth already resident in HBM (CUDA events on the engine stream).
`e2e`     frames/s through the host-buffer C-ABI call lcd_process_frames (pinned host images + depth in,
          word ids + likelihood + verified poses out; H2D/D2H inside the timed region).
`roofline` the dictionary-NN kernel (knn2_tensor_kernel; knn2_hamming_kernel with LCD_NN_TENSOR=0), timed live with CUDA events on its stream.
`cpu_baseline` / --impl reference: cv2.ORB + the oracle port of the reference algorithm on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W_WORDS = 49152
S_SIGS = 10000
F_FEATS = 1000
N_PLACES = 50
IMG_W, IMG_H = 640, 480
DESC_BYTES = 32
NNDR = 0.8
KCAM = (525.0, 525.0, 320.0, 240.0)
METRIC = "loop-closure queries/sec"
UNIT = "queries/s"
STAGES = ["detect(bgr->gray, depth mask, orb detect+describe, 3-D lifting)", "quantise(knn2+nndr+new-words)", "score(tf-idf)",
          "verify(top-1 hypothesis: descriptor matching + pnp-ransac + refinement)"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def tensor_peak_tops():
    """Dense s8 tensor peak in Tera-op/s: MEASURED_PEAKS.json holds the measured bf16 figure (burst: the kernel is timed
    alone); tcgen05 kind::i8 runs at twice the bf16 rate (K=32 per instruction against K=16), so the s8 roof is 2x it."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        if "bf16_tflops" in d:
            return 2.0 * float(d["bf16_tflops"]), "2 x measured bf16 dense (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 2.0 * 2250.0, "2 x nominal bf16 dense (B200_PROFILING.md fallback)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def workload_config(batch: int, where: str):
    return {
        "workload": "BASELINE configs[1]: 640x480 synthetic RGB-D stream, ORB 1000 kp/frame, 49k-word binary dictionary, 10k signatures",
        "words": W_WORDS, "signatures": S_SIGS, "features_per_frame": F_FEATS, "descriptor_bytes": DESC_BYTES, "image": f"{IMG_W}x{IMG_H} BGR8 + depth16",
        "places": N_PLACES, "frames_per_step": batch,
        "mode": "localisation (frozen dictionary + map; per-frame insert/score/roll-back semantics, SURVEY App. C.5)",
        "stages": STAGES,
        "params": "Kp/DetectorStrategy=2 ORB(3 levels, scale 2, edge 19, FAST 20), Kp/MaxFeatures=1000, Mem/DepthAsMask, exact NN (Kp/NNStrategy=0 order), "
                  "Kp/NndrRatio=0.8, Kp/NewWordsComparedTogether, Kp/IncrementalDictionary, Vis/Iterations=300, Vis/PnPReprojError=2, Vis/MinInliers=20, "
                  "Vis/PnPRefineIterations=1, Vis/CorNNDR=0.8; hypothesis = raw-likelihood arg-max",
        "l2": "flushed between timed steps (256 MiB write, outside the timed events)" if where == "gpu" else "n/a",
    }


# ------------------------------------------------------------------------- CPU (reference) arm
def cv2_orb_fn():
    from oracle import feature2d_py as f2d
    from rtabmap_b200 import synth

    p = f2d.OrbParams(n_features=F_FEATS)
    return lambda img, dep: f2d.detect_describe(img, dep, synth.CAMERA_K4, p)


class CpuReference:
    """cv2.ORB (the OpenCV the reference calls) + oracle port of addNewWords / computeLikelihood / computeTransform."""

    def __init__(self, world):
        from oracle import oracle_py as orc

        self.orc = orc
        self.world = world
        self.orb = cv2_orb_fn()
        o = orc.OracleDictionary(0, DESC_BYTES, True, NNDR, True)
        o.add_words(world.word_ids, world.vocab)
        o.last_word_id = int(world.word_ids.max())
        o.update()
        o.load_csr(world.smap.word_ids, world.smap.row_ptr, world.smap.sig, world.smap.cnt)
        o.set_ni(world.smap.sig_ids, world.smap.ni)
        self.o = o

    def one(self, img, dep):
        w = self.world
        kp, d, x = self.orb(img, dep)
        words, like = self.o.localize_ro(d, w.smap.sig_ids, S_SIGS + 1)
        h = int(np.argmax(like))
        n = int(w.smap.ni[h])
        v = self.orc.verify_pair(w.store.desc[h][:n], w.store.xyz[h][:n], d, kp[:, :2], KCAM)
        return kp, d, words, like, int(w.smap.sig_ids[h]), v

    def rate(self, imgs, deps, n_frames: int, threads: int):
        import cv2

        cv2.setNumThreads(1)  # parallelism is over frames
        t0 = time.perf_counter()
        if threads <= 1:
            res = [self.one(imgs[b], deps[b]) for b in range(n_frames)]
        else:
            with ThreadPoolExecutor(threads) as ex:
                res = list(ex.map(lambda b: self.one(imgs[b], deps[b]), range(n_frames)))
        dt = time.perf_counter() - t0
        return n_frames / dt, dt, res


def run_reference(args):
    from rtabmap_b200 import synth

    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    threads = max(1, min(os.cpu_count() or 1, args.ref_threads or (os.cpu_count() or 1)))
    per_step = threads  # one frame per thread and step: a bounded sample of the workload
    world = synth.make_place_world(cv2_orb_fn(), N_PLACES, W_WORDS, S_SIGS, F_FEATS, IMG_H, IMG_W)
    imgs, deps, places = synth.make_view_frames(world, per_step)
    ref = CpuReference(world)
    if args.warmup:
        ref.rate(imgs, deps, min(per_step, 2), threads)
    steps = max(1, min(args.steps, 3))
    total = 0.0
    for _ in range(steps):
        _, dt, res = ref.rate(imgs, deps, per_step, threads)
        total += dt
    value = per_step * steps / total
    hit = float(np.mean([world.sig_place[r[4] - 1] == places[b] for b, r in enumerate(res)]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(per_step, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} frames/step x {steps} steps of the same workload: cv2.ORB (OpenCV 4.13, 1 thread per frame) + oracle port of "
                                   f"VWDictionary::addNewWords, Memory::computeLikelihood (std::map structures as in the reference) and "
                                   f"Memory::computeTransform; frames spread over {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "top1_place_hit_rate": hit, "verified_rate": float(np.mean([r[5]["ok"] for r in res])),
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from rtabmap_b200 import Engine, sharding, synth

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world_size}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world_size > 1:
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # weak scaling: every GPU brings its own --batch frames per step (a relocalisation service adds cameras with GPUs); the
    # dictionary and the inverted index are sharded by word range, so each rank still searches ALL frames' descriptors.
    BL = args.batch                # frames per step detected and verified by this rank
    B = BL * world_size            # frames per step of the whole job
    n_pool = 3

    eng = Engine(device=local, desc_dim=DESC_BYTES, max_words=W_WORDS, max_signatures=S_SIGS + 2, max_queries=F_FEATS, max_batch=B)
    op = Engine.orb_params(KCAM, n_features=F_FEATS)
    vp = Engine.verify_params(KCAM)

    def gpu_orb_fn(img, dep):
        return eng.orb_detect_describe(img[None], dep[None], op, cap=F_FEATS)[0]

    t0 = time.time()
    world = synth.make_place_world(gpu_orb_fn, N_PLACES, W_WORDS, S_SIGS, F_FEATS, IMG_H, IMG_W)
    imgs_all, deps_all, places = synth.make_view_frames(world, BL * n_pool, seed=3 + rank)  # this rank's frames only
    log(f"[rank {rank}] world built in {time.time() - t0:.1f}s: {len(world.vocab)} words, {world.smap.nnz} postings")

    r0, r1 = sharding.shard_rows(W_WORDS, world_size, rank)
    ids, vocab, smap = world.word_ids, world.vocab, world.smap
    eng.add_words(ids[r0:r1], vocab[r0:r1])
    eng.last_word_id = W_WORDS
    eng.update()
    if world_size > 1:
        eng.shard_set_row_offset(r0)
        w_, p_, s_, c_ = sharding.shard_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt, ids[r0:r1])
        eng.load_csr(w_, p_, s_, c_)
    else:
        eng.load_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt)
    eng.set_ni(smap.sig_ids, smap.ni)
    for s0 in range(0, S_SIGS, 1000):  # signature store, replicated on every rank
        eng.sig_add_batch(smap.sig_ids[s0:s0 + 1000], world.store.desc[s0:s0 + 1000], world.store.xyz[s0:s0 + 1000], smap.ni[s0:s0 + 1000])

    ext = torch.cuda.ExternalStream(eng.stream, device=local)
    torch.cuda.set_stream(ext)
    nq = B * F_FEATS
    f0, f1 = rank * BL, (rank + 1) * BL  # this rank's rows in the job-wide (all-gathered) arrays
    nf = BL
    d_img = [torch.from_numpy(imgs_all[k * BL:(k + 1) * BL]).cuda() for k in range(n_pool)]
    d_dep = [torch.from_numpy(deps_all[k * BL:(k + 1) * BL].view(np.int16)).cuda() for k in range(n_pool)]
    h_img = [torch.from_numpy(imgs_all[k * BL:(k + 1) * BL]).pin_memory() for k in range(n_pool)]
    h_dep = [torch.from_numpy(deps_all[k * BL:(k + 1) * BL].view(np.int16)).pin_memory() for k in range(n_pool)]
    d_sig = torch.from_numpy(smap.sig_ids).cuda()
    h_sig = torch.from_numpy(smap.sig_ids).pin_memory()
    d_words = torch.zeros(nq, dtype=torch.int32, device="cuda")
    d_like = torch.zeros(B * S_SIGS, dtype=torch.float32, device="cuda")
    h_words = torch.zeros((B, F_FEATS), dtype=torch.int32).pin_memory()
    h_like = torch.zeros((B, S_SIGS), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    img_bytes = IMG_W * IMG_H * 3
    dep_bytes = IMG_W * IMG_H * 2
    if world_size > 1:
        d_desc_loc = torch.zeros(nf * F_FEATS * DESC_BYTES, dtype=torch.uint8, device="cuda")
        d_uv_loc = torch.zeros(nf * F_FEATS * 2, dtype=torch.float32, device="cuda")
        d_n_loc = torch.zeros(nf, dtype=torch.int32, device="cuda")
        d_desc = torch.zeros(nq * DESC_BYTES, dtype=torch.uint8, device="cuda")
        d_uv = torch.zeros(nq * 2, dtype=torch.float32, device="cuda")
        d_keys = torch.zeros(nq * 2, dtype=torch.int32, device="cuda")
        d_keys_all = torch.zeros(world_size * nq * 2, dtype=torch.int32, device="cuda")
        d_rowids = torch.from_numpy(ids).cuda()
        d_scores = torch.zeros(B * S_SIGS, dtype=torch.int64, device="cuda")
        d_n_all = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_scores_loc = torch.zeros(nf * S_SIGS, dtype=torch.int64, device="cuda")
        d_words_loc = torch.zeros(nf * F_FEATS, dtype=torch.int32, device="cuda")

    def sharded_step(img_t, dep_t):
        # detect: frames sharded; quantise: words sharded; score: words sharded; verify: frames sharded
        d_desc_loc.zero_()
        eng._check(eng._lib.lcd_orb_detect_describe_dev(eng.handle, nf, ctypes.c_void_p(img_t.data_ptr()), IMG_W, IMG_H, 3,
                                                         ctypes.c_void_p(dep_t.data_ptr()), 1, ctypes.byref(op), F_FEATS, None,
                                                         ctypes.c_void_p(d_desc_loc.data_ptr()), None, ctypes.c_void_p(d_uv_loc.data_ptr()),
                                                         ctypes.c_void_p(d_n_loc.data_ptr()), None))
        dist.all_gather_into_tensor(d_desc, d_desc_loc)
        dist.all_gather_into_tensor(d_uv, d_uv_loc)
        dist.all_gather_into_tensor(d_n_all, d_n_loc)
        eng.shard_knn2_keys_dev(d_desc.data_ptr(), nq, d_keys.data_ptr())
        dist.all_gather_into_tensor(d_keys_all, d_keys)
        # stage 2 sharded by frame: resolve the local frames, all-gather their word ids, score every frame on the local word range
        eng.shard_resolve_frames_dev(d_desc.data_ptr(), f0, nf, B, F_FEATS, d_keys_all.data_ptr(), world_size, d_rowids.data_ptr(), W_WORDS,
                                     d_n_all.data_ptr(), d_words_loc.data_ptr(), True, NNDR, True)
        dist.all_gather_into_tensor(d_words, d_words_loc)
        eng.shard_score_ids_dev(d_words.data_ptr(), B, F_FEATS, d_sig.data_ptr(), S_SIGS, S_SIGS + 1, d_scores.data_ptr())
        # every rank only needs the likelihood rows of the frames it verifies: reduce-scatter (half the traffic of an all-reduce);
        # the int64 fixed-point sums are exact, so the result does not depend on the reduction order
        dist.reduce_scatter_tensor(d_scores_loc, d_scores, op=dist.ReduceOp.SUM)
        eng.shard_finalize_dev(d_scores_loc.data_ptr(), nf * S_SIGS, d_like.data_ptr() + f0 * S_SIGS * 4)
        eng.verify_top_dev(d_desc.data_ptr() + f0 * F_FEATS * DESC_BYTES, d_uv.data_ptr() + f0 * F_FEATS * 8, nf, F_FEATS,
                           d_like.data_ptr() + f0 * S_SIGS * 4, d_sig.data_ptr(), S_SIGS, vp)

    def step_dev(k):
        if world_size == 1:
            eng.process_frames_dev(d_img[k % n_pool].data_ptr(), B, IMG_W, IMG_H, 3, d_dep[k % n_pool].data_ptr(), 1, op, d_sig.data_ptr(), S_SIGS,
                                   S_SIGS + 1, vp, d_words.data_ptr(), d_like.data_ptr(), True, NNDR, True)
        else:
            sharded_step(d_img[k % n_pool], d_dep[k % n_pool])

    def step_host(k):
        hi, hd = h_img[k % n_pool], h_dep[k % n_pool]
        if world_size == 1:
            nkp, _, _, hyp, res = eng.process_frames(hi.numpy(), hd.numpy().view(np.uint16), op, h_sig.numpy(), S_SIGS + 1, vp, True, NNDR, True,
                                                     out_words=h_words.numpy(), out_like=h_like.numpy())
            return hyp, res
        di, dd = d_img[0], d_dep[0]
        di.copy_(hi, non_blocking=True)  # every rank uploads the frames it detects
        dd.copy_(hd, non_blocking=True)
        sharded_step(di, dd)
        h_words.view(-1).copy_(d_words, non_blocking=True)
        h_like.view(-1).copy_(d_like, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return eng.process_fetch(nf)

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) -------------------------------------------------
    for k in range(args.warmup):
        step_dev(k)
    barrier()
    eng.profile_enable(True)
    eng.profile_reset()
    launches0 = eng.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    torch.cuda.cudart().cudaProfilerStart()  # lets `ncu --profile-from-start off` see exactly the timed region
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush, outside the timed events
        ev[k][0].record(ext)
        step_dev(k)
        ev[k][1].record(ext)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    torch.cuda.cudart().cudaProfilerStop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launch_count - launches0
    prof = {name: eng.profile_read(i) for i, name in enumerate(["nn", "resolve", "score", "match", "pnp", "orb"])}
    eng.profile_enable(False)
    if world_size > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = B * args.steps / (dev_ms * 1e-3)

    # sanity inside the bench: the verified hypothesis must be a view of the revisited place
    last_pool = (args.steps - 1) % n_pool
    true_places = places[last_pool * BL:(last_pool + 1) * BL]
    hyp_d, res_d = eng.process_fetch(nf)
    hit = float(np.mean(world.sig_place[np.maximum(hyp_d, 1) - 1] == true_places))
    verified = float(np.mean([r["ok"] for r in res_d]))

    # ---- end-to-end timing through the host-buffer C ABI ------------------------------------
    for k in range(min(args.warmup, 3)):
        step_host(k)
    barrier()
    e2e_s = 0.0
    if world_size == 1:
        # the call a relocalisation service makes: lcd_process_frames_submit / _wait, two batches in flight, so the PCIe upload
        # of batch k+1 runs under the kernels of batch k.  Every step's images + depth come from pinned host memory and every
        # step's word ids, likelihood rows, hypotheses and verification results are copied back; all of it inside the timed
        # region, and so is the L2 flush between steps (queued on the engine stream).
        from rtabmap_b200.capi import VerifyResult
        sets = []
        for _ in range(2):
            sets.append(dict(nkp=torch.zeros(B, dtype=torch.int32).pin_memory(), words=torch.zeros((B, F_FEATS), dtype=torch.int32).pin_memory(),
                             like=torch.zeros((B, S_SIGS), dtype=torch.float32).pin_memory(), hyp=torch.zeros(B, dtype=torch.int32).pin_memory(),
                             res=(VerifyResult * B)()))
        sig_np = h_sig.numpy()

        def submit(k):
            o = sets[k & 1]
            eng.process_frames_submit(h_img[k % n_pool].numpy(), h_dep[k % n_pool].numpy().view(np.uint16), op, sig_np, S_SIGS + 1, vp,
                                      o["nkp"].numpy(), o["words"].numpy(), o["like"].numpy(), o["hyp"].numpy(), o["res"], True, NNDR, True)

        def collect(k):
            eng.process_frames_wait()
            o = sets[k & 1]
            return o["hyp"].numpy().copy(), [{"ok": int(r.ok)} for r in o["res"]]

        for k in range(2):  # warm-up of the pipelined path (allocates the two staging slots)
            submit(k)
        collect(0)
        collect(1)
        torch.cuda.synchronize()
        with torch.cuda.stream(ext):
            t0 = time.perf_counter()
            trace = [] if os.environ.get("LCD_BENCH_TRACE") else None
            submit(0)
            for k in range(1, args.steps):
                flush.fill_(k & 0xFF)
                ta = time.perf_counter()
                submit(k)
                tb = time.perf_counter()
                hyp_h, res_h = collect(k - 1)
                if trace is not None:
                    trace.append((round((ta - t0) * 1e3, 2), round((tb - ta) * 1e3, 2), round((time.perf_counter() - tb) * 1e3, 2)))
            hyp_h, res_h = collect(args.steps - 1)
            e2e_s = time.perf_counter() - t0
            if trace:
                print("e2e trace (t_submit_ms, submit_call_ms, collect_call_ms):", trace[:10], file=sys.stderr)
    else:
        # sharded job: the public calls are the *_dev entry points, so the caller owns the copies.  Double-buffered: the upload of
        # step k+1 (this rank's frames, pinned host memory) runs on a side stream under the kernels and collectives of step k;
        # every step's word ids, likelihood rows, hypotheses and verification results are copied back inside the timed region.
        up = torch.cuda.Stream()
        up_ev = [torch.cuda.Event(), torch.cuda.Event()]

        def prefetch(k):
            with torch.cuda.stream(up):
                d_img[k & 1].copy_(h_img[k % n_pool], non_blocking=True)
                d_dep[k & 1].copy_(h_dep[k % n_pool], non_blocking=True)
                up_ev[k & 1].record(up)

        barrier()
        t0 = time.perf_counter()
        prefetch(0)
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            ext.wait_event(up_ev[k & 1])
            sharded_step(d_img[k & 1], d_dep[k & 1])
            h_words.view(-1)[f0 * F_FEATS:f1 * F_FEATS].copy_(d_words[f0 * F_FEATS:f1 * F_FEATS], non_blocking=True)
            h_like.view(-1)[f0 * S_SIGS:f1 * S_SIGS].copy_(d_like[f0 * S_SIGS:f1 * S_SIGS], non_blocking=True)
            if k + 1 < args.steps:
                prefetch(k + 1)
            hyp_h, res_h = eng.process_fetch(nf)  # device-wide synchronisation + results of this rank's frames
        e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    if world_size > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B * args.steps / e2e_s
    e2e_hit = float(np.mean(world.sig_place[np.maximum(hyp_h, 1) - 1] == true_places))
    e2e_verified = float(np.mean([r["ok"] for r in res_h]))

    if rank != 0:
        if world_size > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src, sm_max = peaks()
    rows_local = r1 - r0
    alg_bytes = rows_local * DESC_BYTES + nq * DESC_BYTES + nq * 16  # W*D + Q*D + Q*16 per launch
    nn_ms, nn_launches = prof["nn"]
    nn_avg_s = (nn_ms / max(nn_launches, 1)) * 1e-3
    achieved = alg_bytes / nn_avg_s / 1e9 if nn_avg_s > 0 else 0.0
    pairs = float(rows_local) * nq
    popc_per_pair = 5  # LCD_NN_VARIANT=2: partial carry-save tree, 8 XOR words -> 5 POPC
    sm_clk = (clocks or {}).get("sm_mhz") or sm_max
    popc_peak = 16.0 * 148 * sm_clk * 1e6  # 16 POPC lanes / clk / SM, measured (profiles/r01_nn_sweep.json)
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("knn2_hamming_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    share = dict({k + "_ms": v[0] / args.steps for k, v in prof.items()}, step_ms=dev_ms / args.steps)
    if os.environ.get("LCD_NN_TENSOR", "1") != "0":
        # dominant kernel: knn2_tensor_kernel (tcgen05 kind::i8).  Algorithmic work per launch = one s8 multiply-add per
        # (query, word, descriptor bit): 2 * Q * W * 256 operations.
        ops = 2.0 * pairs * DESC_BYTES * 8
        tpeak, tsrc = tensor_peak_tops()
        ach = ops / nn_avg_s / 1e12 if nn_avg_s > 0 else 0.0
        traffic_t = None
        if tp.exists():
            try:
                traffic_t = json.loads(tp.read_text()).get("knn2_tensor_kernel_dram_bytes_per_launch")
            except Exception:
                traffic_t = None
        roofline = {
            "kernel": "knn2_tensor_kernel (tcgen05.mma kind::i8, M128 N256 K32)", "bound": "tensor", "achieved": ach, "peak": tpeak,
            "unit": "TFLOP/s", "frac": ach / tpeak, "peak_source": tsrc, "traffic": traffic_t,
            "algorithmic_ops_per_launch": ops, "algorithmic_bytes_per_launch": alg_bytes,
            "avg_launch_ms": nn_avg_s * 1e3, "launches_timed": int(nn_launches),
            "note": "s8 operations counted as FLOPs of the +-1 encoded Hamming GEMM; results are exact integers (parity tests)",
            "pairs_per_s": pairs / nn_avg_s if nn_avg_s > 0 else 0.0, "step_share_ms": share,
        }
    else:
        roofline = {
            "kernel": "knn2_hamming_kernel<8,8,2>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
            "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
            "avg_launch_ms": nn_avg_s * 1e3, "launches_timed": int(nn_launches),
            "binding_resource": "integer pipes (POPC 16 lanes/clk/SM on XU + LOP3 on ALU): the vocabulary is SMEM/L2 resident, see DESIGN.md §4",
            "pairs_per_s": pairs / nn_avg_s if nn_avg_s > 0 else 0.0,
            "popc_per_s": pairs * popc_per_pair / nn_avg_s if nn_avg_s > 0 else 0.0, "popc_peak_per_s": popc_peak,
            "popc_frac": (pairs * popc_per_pair / nn_avg_s) / popc_peak if nn_avg_s > 0 else 0.0,
            "step_share_ms": share,
        }

    # the XOR/POPC kernel on the same resident data (three untimed-region steps), for comparison with the tensor kernel
    if world_size == 1 and os.environ.get("LCD_NN_TENSOR", "1") != "0":
        eng.nn_select(0)
        eng.profile_enable(True)
        eng.profile_reset()
        for k in range(3):
            step_dev(k)
        torch.cuda.synchronize()
        p_ms, p_n = eng.profile_read(0)
        eng.profile_enable(False)
        eng.nn_select(1)
        p_s = (p_ms / max(p_n, 1)) * 1e-3
        roofline["popcount_kernel"] = {
            "kernel": "knn2_hamming_kernel<8,8,2> (lcd_nn_select(e, 0))", "avg_launch_ms": p_s * 1e3, "launches_timed": int(p_n),
            "bound": "integer pipes (POPC 16 lanes/clk/SM)", "popc_per_s": pairs * popc_per_pair / p_s if p_s > 0 else 0.0,
            "popc_peak_per_s": popc_peak, "popc_frac": (pairs * popc_per_pair / p_s) / popc_peak if p_s > 0 else 0.0,
            "hbm_GBps": alg_bytes / p_s / 1e9 if p_s > 0 else 0.0, "hbm_frac": (alg_bytes / p_s / 1e9) / hbm_peak if p_s > 0 else 0.0,
            "traffic": traffic,
        }

    cpu = None
    if world_size == 1 and not args.no_cpu_baseline:
        threads = max(1, min(os.cpu_count() or 1, 32))
        ref = CpuReference(world)
        rate, dt, res = ref.rate(imgs_all, deps_all, threads, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{threads} frames of the same workload ({dt:.1f} s wall): cv2.ORB (OpenCV 4.13) + oracle port (std::map inverted index, scalar "
                         f"popcount NN, restated EPnP/RANSAC), {threads} threads, one frame per thread"}
        # cross-check while we are here: the first frames against the GPU result (the parity tests proper are tests/)
        nchk = min(2, threads)
        nkp, w_g, l_g, hyp_c, res_c = eng.process_frames(imgs_all[:nchk], deps_all[:nchk], op, smap.sig_ids, S_SIGS + 1, vp, True, NNDR, True)
        for b in range(nchk):
            kp, d, words, like, hyp_o, v = res[b]
            assert nkp[b] == len(kp) and np.array_equal(w_g[b][:len(words)], words), "GPU/oracle word ids differ"
            assert np.allclose(l_g[b], like, atol=1e-4, rtol=1e-4), "GPU/oracle likelihood differ"
            assert hyp_o == hyp_c[b] and v["ok"] == res_c[b]["ok"] and len(v["inliers"]) == res_c[b]["n_inliers"], "GPU/oracle verification differ"
            assert np.allclose(v["rvec"], res_c[b]["rvec"], atol=1e-4) and np.allclose(v["tvec"], res_c[b]["tvec"], atol=1e-4)

    par = "single GPU" if world_size == 1 else (f"x{world_size}: every GPU detects and verifies its own {BL} frames per step, dictionary + inverted index sharded by word range; "
                                                  "all-gather(descriptors, top-2 keys, word ids) + reduce-scatter(int64 scores) over NCCL")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": dict(workload_config(B, "gpu"), frames_per_gpu=BL, parallelism=par),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(B * (img_bytes + dep_bytes) + S_SIGS * 4),
                "d2h_bytes_per_step": int(nq * 4 + B * S_SIGS * 4 + B * (4 + 4 + 124)),
                "api": "lcd_process_frames_submit/_wait (pinned host buffers, 2 batches in flight, L2 flush between steps inside the timed region)" if world_size == 1 else "sharded *_dev calls, double-buffered pinned uploads on a side stream, results copied back every step",
                "top1_place_hit_rate": e2e_hit, "verified_rate": e2e_verified},
        "roofline": roofline, "cpu_baseline": cpu, "top1_place_hit_rate": hit, "verified_rate": verified, "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="frames per step")
    ap.add_argument("--ref-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
