#!/usr/bin/env python
"""bench.py — loop-closure queries/sec on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference          # the reference's CPU algorithm on the host cores

One "step" = one batch of B independent localisation queries (frames) through the hot path that is
built so far (see `config.stages`): each frame = F=1000 ORB-sized binary descriptors quantised against
the W=49 152-word dictionary (exact 2-NN + NNDR + intra-frame new words, VWDictionary::addNewWords),
scored by TF-IDF over the S=10 000-signature inverted index (Memory::computeLikelihood), and the top
hypothesis verified geometrically (Memory::computeTransform: descriptor matching + PnP RANSAC + refinement).
Workload = BASELINE.json configs[1] (640x480 stream, ORB 1000 kp/frame, 49k words, 10k signatures).

`value`  : frames/s with the descriptors already resident in HBM (CUDA events on the engine stream).
`e2e`    : frames/s through the host-buffer C-ABI call lcd_process_batch (pinned host descriptors + keypoints in,
           word ids + likelihood vectors + verified poses out, H2D/D2H inside the timed region).
`roofline`: the dictionary-NN kernel (knn2_hamming_kernel), timed live with CUDA events on its stream.
`cpu_baseline`: the oracle port of the reference algorithm on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W_WORDS = 49152
S_SIGS = 10000
F_FEATS = 1000
DESC_BYTES = 32
NNDR = 0.8
KCAM = (525.0, 525.0, 320.0, 240.0)
METRIC = "loop-closure queries/sec"
UNIT = "queries/s"
STAGES_BUILT = ["quantise(knn2+nndr+new-words)", "score(tf-idf)", "verify(top-1 hypothesis: descriptor matching + pnp-ransac + refinement)"]
STAGES_MISSING = ["detect(orb)"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(batch: int, n_batches: int):
    from rtabmap_b200 import synth

    vocab = synth.make_binary_vocabulary(W_WORDS, DESC_BYTES, seed=1)
    ids = np.arange(1, W_WORDS + 1, dtype=np.int32)
    smap = synth.make_map(ids, S_SIGS, F_FEATS, seed=2)
    store = synth.make_signature_store(vocab, ids, smap, seed=4)
    q, uv, places, poses = synth.make_query_frames_geo(store, smap, batch * n_batches, seed=3)
    return vocab, ids, smap, store, q, uv, places


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


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


# ------------------------------------------------------------------------- CPU (reference) arm
def cpu_reference_rate(vocab, ids, smap, store, frames, uv, n_frames: int, threads: int):
    """Oracle port of addNewWords + computeLikelihood (roll-back semantics) + computeTransform of the top hypothesis."""
    from oracle import oracle_py as orc

    o = orc.OracleDictionary(0, DESC_BYTES, True, NNDR, True)
    o.add_words(ids, vocab)
    o.last_word_id = int(ids.max())
    o.update()
    o.load_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt)

    def one(b):
        fq = frames[b * F_FEATS:(b + 1) * F_FEATS]
        w, l = o.localize_ro(fq, smap.sig_ids, S_SIGS + 1)
        h = int(np.argmax(l))
        v = orc.verify_pair(store.desc[h], store.xyz[h], fq, uv[b * F_FEATS:(b + 1) * F_FEATS], KCAM)
        return w, l, int(smap.sig_ids[h]), v

    t0 = time.perf_counter()
    if threads <= 1:
        res = [one(b) for b in range(n_frames)]
    else:
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(one, range(n_frames)))
    dt = time.perf_counter() - t0
    return n_frames / dt, dt, res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = max(1, min(os.cpu_count() or 1, args.ref_threads or (os.cpu_count() or 1)))
    per_step = threads  # one frame per thread and step: a bounded sample of the workload
    vocab, ids, smap, store, q, uv, places = make_workload(per_step, 1)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_rate(vocab, ids, smap, store, q, uv, min(per_step, 2), threads)
    steps = max(1, min(args.steps, 3))
    times = []
    for _ in range(steps):
        rate, dt, _ = cpu_reference_rate(vocab, ids, smap, store, q, uv, per_step, threads)
        times.append(dt)
    total = float(sum(times))
    value = per_step * steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1),
        "ms_per_step": 1e3 * total / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": workload_config(per_step, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} frames/step x {steps} steps of the same workload, oracle port of VWDictionary::addNewWords + "
                                   f"Memory::computeLikelihood (std::map structures as in the reference) + Memory::computeTransform of the top hypothesis, "
                                   f"frames spread over {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(batch: int, where: str):
    return {
        "workload": "BASELINE configs[1]: 640x480 synthetic RGB-D stream, ORB 1000 kp/frame, 49k-word binary dictionary, 10k signatures",
        "words": W_WORDS, "signatures": S_SIGS, "features_per_frame": F_FEATS, "descriptor_bytes": DESC_BYTES,
        "frames_per_step": batch, "mode": "localisation (frozen dictionary + map; per-frame insert/score/roll-back semantics, SURVEY App. C.5)",
        "stages": STAGES_BUILT, "stages_not_yet_in_step": STAGES_MISSING,
        "nn": "exact (Kp/NNStrategy=0 order), Kp/NndrRatio=0.8, Kp/NewWordsComparedTogether=true, Kp/IncrementalDictionary=true",
        "l2": "flushed between timed steps (256 MiB write, outside the timed events)" if where == "gpu" else "n/a",
    }


# ------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist

    from rtabmap_b200 import Engine, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl b200) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    B = args.batch
    n_pool = 4
    vocab, ids, smap, store, q_all, uv_all, places = make_workload(B, n_pool)
    r0, r1 = sharding.shard_rows(W_WORDS, world, rank)
    eng = Engine(device=local, desc_dim=DESC_BYTES, max_words=W_WORDS, max_signatures=S_SIGS + 2, max_queries=F_FEATS, max_batch=B)
    eng.add_words(ids[r0:r1], vocab[r0:r1])
    eng.last_word_id = W_WORDS
    eng.update()
    if world > 1:
        eng.shard_set_row_offset(r0)
        w_, p_, s_, c_ = sharding.shard_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt, ids[r0:r1])
        eng.load_csr(w_, p_, s_, c_)
        eng.set_ni(smap.sig_ids, smap.ni)
    else:
        eng.load_csr(smap.word_ids, smap.row_ptr, smap.sig, smap.cnt)
    for s0 in range(0, S_SIGS, 1000):  # signature store (descriptors + 3-D points of every node), replicated on every rank
        eng.sig_add_batch(smap.sig_ids[s0:s0 + 1000], store.desc[s0:s0 + 1000], store.xyz[s0:s0 + 1000])
    vp = Engine.verify_params(KCAM)

    ext = torch.cuda.ExternalStream(eng.stream, device=local)
    torch.cuda.set_stream(ext)
    nq = B * F_FEATS
    d_q = [torch.from_numpy(q_all[k * nq:(k + 1) * nq]).cuda() for k in range(n_pool)]
    d_uv = [torch.from_numpy(uv_all[k * nq:(k + 1) * nq]).cuda() for k in range(n_pool)]
    h_uv = [torch.from_numpy(uv_all[k * nq:(k + 1) * nq]).pin_memory() for k in range(n_pool)]
    d_sig = torch.from_numpy(smap.sig_ids).cuda()
    d_words = torch.zeros(nq, dtype=torch.int32, device="cuda")
    d_like = torch.zeros(B * S_SIGS, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if world > 1:
        d_keys = torch.zeros(nq * 2, dtype=torch.int32, device="cuda")
        d_keys_all = torch.zeros(world * nq * 2, dtype=torch.int32, device="cuda")
        d_rowids = torch.from_numpy(ids).cuda()
        d_scores = torch.zeros(B * S_SIGS, dtype=torch.int64, device="cuda")
    h_q = [torch.from_numpy(q_all[k * nq:(k + 1) * nq]).pin_memory() for k in range(n_pool)]
    h_words = torch.zeros((B, F_FEATS), dtype=torch.int32).pin_memory()
    h_like = torch.zeros((B, S_SIGS), dtype=torch.float32).pin_memory()
    h_sig = torch.from_numpy(smap.sig_ids).pin_memory()

    def step_dev(k):
        dq = d_q[k % n_pool]
        du = d_uv[k % n_pool]
        if world == 1:
            eng.process_batch_dev(dq.data_ptr(), du.data_ptr(), B, F_FEATS, d_sig.data_ptr(), S_SIGS, S_SIGS + 1, vp, d_words.data_ptr(),
                                  d_like.data_ptr(), True, NNDR, True)
        else:
            eng.shard_knn2_keys_dev(dq.data_ptr(), nq, d_keys.data_ptr())
            dist.all_gather_into_tensor(d_keys_all, d_keys)
            eng.shard_resolve_score_dev(dq.data_ptr(), B, F_FEATS, d_keys_all.data_ptr(), world, d_rowids.data_ptr(), W_WORDS, W_WORDS,
                                        d_sig.data_ptr(), S_SIGS, S_SIGS + 1, d_words.data_ptr(), d_scores.data_ptr(), True, NNDR, True)
            dist.all_reduce(d_scores, op=dist.ReduceOp.SUM)
            eng.shard_finalize_dev(d_scores.data_ptr(), B * S_SIGS, d_like.data_ptr())
            # every rank verifies its share of the frames against the replicated signature store
            eng.verify_top_dev(dq.data_ptr() + f0 * F_FEATS * DESC_BYTES, du.data_ptr() + f0 * F_FEATS * 8, f1 - f0, F_FEATS,
                               d_like.data_ptr() + f0 * S_SIGS * 4, d_sig.data_ptr(), S_SIGS, vp)

    def step_host(k):
        hq = h_q[k % n_pool]
        hu = h_uv[k % n_pool]
        if world == 1:
            _, _, hyp, res = eng.process_batch(hq.numpy(), hu.numpy(), B, h_sig.numpy(), S_SIGS + 1, vp, True, NNDR, True,
                                               out_words=h_words.numpy(), out_like=h_like.numpy())
            return hyp, res
        else:
            dq = d_q[0]
            du = d_uv[0]
            dq.copy_(hq, non_blocking=True)
            du.copy_(hu, non_blocking=True)
            eng.shard_knn2_keys_dev(dq.data_ptr(), nq, d_keys.data_ptr())
            dist.all_gather_into_tensor(d_keys_all, d_keys)
            eng.shard_resolve_score_dev(dq.data_ptr(), B, F_FEATS, d_keys_all.data_ptr(), world, d_rowids.data_ptr(), W_WORDS, W_WORDS,
                                        d_sig.data_ptr(), S_SIGS, S_SIGS + 1, d_words.data_ptr(), d_scores.data_ptr(), True, NNDR, True)
            dist.all_reduce(d_scores, op=dist.ReduceOp.SUM)
            eng.shard_finalize_dev(d_scores.data_ptr(), B * S_SIGS, d_like.data_ptr())
            eng.verify_top_dev(dq.data_ptr() + f0 * F_FEATS * DESC_BYTES, du.data_ptr() + f0 * F_FEATS * 8, f1 - f0, F_FEATS,
                               d_like.data_ptr() + f0 * S_SIGS * 4, d_sig.data_ptr(), S_SIGS, vp)
            h_words.view(-1).copy_(d_words, non_blocking=True)
            h_like.view(-1).copy_(d_like, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return eng.process_fetch(f1 - f0)

    f0, f1 = sharding.shard_rows(B, world, rank)  # frames this rank verifies

    def barrier():
        if world > 1:
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
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush, outside the timed events
        ev[k][0].record(ext)
        step_dev(k)
        ev[k][1].record(ext)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = eng.launch_count - launches0
    nn_ms, nn_launches = eng.profile_read(0)
    res_ms, _ = eng.profile_read(1)
    sc_ms, _ = eng.profile_read(2)
    mt_ms, _ = eng.profile_read(3)
    pnp_ms, _ = eng.profile_read(4)
    eng.profile_enable(False)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = B * args.steps / (dev_ms * 1e-3)

    # sanity inside the bench: the revisited place must win the likelihood of its frame
    torch.cuda.synchronize()
    last_pool = (args.steps - 1) % n_pool
    best = d_like.view(B, S_SIGS).argmax(dim=1).cpu().numpy()
    hit = float(np.mean(smap.sig_ids[best] == places[last_pool * B:(last_pool + 1) * B]))
    hyp_d, res_d = eng.process_fetch(f1 - f0)
    verified = float(np.mean([r["ok"] for r in res_d]))
    assert np.array_equal(hyp_d, places[last_pool * B + f0:last_pool * B + f1]), "verified hypothesis is not the revisited place"

    # ---- end-to-end timing through the host-buffer C ABI ------------------------------------
    for k in range(min(args.warmup, 3)):
        step_host(k)
    barrier()
    e2e_s = 0.0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hyp_h, res_h = step_host(k)
        e2e_s += time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = B * args.steps / e2e_s
    e2e_hit = float(np.mean(smap.sig_ids[h_like.numpy().argmax(axis=1)] == places[last_pool * B:(last_pool + 1) * B]))
    e2e_verified = float(np.mean([r["ok"] for r in res_h]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, peak_src, sm_max = peaks()
    rows_local = r1 - r0
    alg_bytes = rows_local * DESC_BYTES + nq * DESC_BYTES + nq * 16  # W*D + Q*D + Q*16 per launch
    nn_avg_s = (nn_ms / max(nn_launches, 1)) * 1e-3
    achieved = alg_bytes / nn_avg_s / 1e9 if nn_avg_s > 0 else 0.0
    pairs = float(rows_local) * nq
    popc_per_pair = 5  # LCD_NN_VARIANT=2: partial carry-save tree, 8 XOR words -> 5 POPC
    sm_clk = (clocks or {}).get("sm_mhz") or sm_max
    popc_peak = 16.0 * 148 * sm_clk * 1e6  # measured: 16 POPC lanes / clk / SM (profiles/r01_nn_sweep.md)
    traffic = None
    tp = ROOT / "profiles" / "roofline_traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get("knn2_hamming_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "kernel": "knn2_hamming_kernel<8,8,2>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes,
        "avg_launch_ms": nn_avg_s * 1e3, "launches_timed": int(nn_launches),
        "binding_resource": "integer pipe (POPC 16 lanes/clk/SM + LOP3): vocabulary is L2-resident, see DESIGN.md",
        "pairs_per_s": pairs / nn_avg_s if nn_avg_s > 0 else 0.0,
        "popc_per_s": pairs * popc_per_pair / nn_avg_s if nn_avg_s > 0 else 0.0, "popc_peak_per_s": popc_peak,
        "popc_frac": (pairs * popc_per_pair / nn_avg_s) / popc_peak if nn_avg_s > 0 else 0.0,
        "step_share": {"nn_ms": nn_ms / args.steps, "resolve_ms": res_ms / args.steps, "score_ms": sc_ms / args.steps, "match_ms": mt_ms / args.steps, "pnp_ms": pnp_ms / args.steps, "step_ms": dev_ms / args.steps},
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = max(1, min(os.cpu_count() or 1, 32))
        n_cpu = threads
        rate, dt, res = cpu_reference_rate(vocab, ids, smap, store, q_all, uv_all, n_cpu, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n_cpu} frames of the same workload ({dt:.1f} s wall), oracle port (std::map inverted index, scalar popcount NN, restated EPnP/RANSAC), "
                         f"{threads} threads, one frame per thread"}
        # cross-check while we are here: first frames of pool 0 against the GPU result
        nchk = min(2, n_cpu)
        _, _, hyp_c, res_c = eng.process_batch(q_all[:F_FEATS * nchk], uv_all[:F_FEATS * nchk], nchk, smap.sig_ids, S_SIGS + 1, vp, True, NNDR, True,
                                               out_words=h_words.numpy()[:nchk], out_like=h_like.numpy()[:nchk])
        for b in range(nchk):
            assert np.array_equal(res[b][0], h_words.numpy()[b]), "GPU/oracle word ids differ"
            assert np.allclose(res[b][1], h_like.numpy()[b], atol=1e-4, rtol=1e-4), "GPU/oracle likelihood differ"
            assert res[b][2] == hyp_c[b] and res[b][3]["ok"] == res_c[b]["ok"] and len(res[b][3]["inliers"]) == res_c[b]["n_inliers"]
            assert np.allclose(res[b][3]["rvec"], res_c[b]["rvec"], atol=1e-4) and np.allclose(res[b][3]["tvec"], res_c[b]["tvec"], atol=1e-4)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": dict(workload_config(B, "gpu"), parallelism=("single GPU" if world == 1 else f"word-range shards x{world}: all-gather(top-2 keys) + all-reduce(int64 scores)")),
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nq * DESC_BYTES + nq * 8 + S_SIGS * 4),
                "d2h_bytes_per_step": int(nq * 4 + B * S_SIGS * 4 + B * (4 + 124)), "api": "lcd_process_batch (host buffers)" if world == 1 else "sharded *_dev calls + pinned copies",
                "top1_place_hit_rate": e2e_hit, "verified_rate": e2e_verified},
        "roofline": roofline, "cpu_baseline": cpu, "top1_place_hit_rate": hit, "verified_rate": verified, "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="frames per step")
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
