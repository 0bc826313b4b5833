"""Seeded synthetic workloads for the quantise -> score path.

Input generation only (numpy, host): vocabularies, maps (inverted index in CSR form) and query
frames of the shapes BASELINE.json names.  Seeds: vocabulary 1, map 2, stream 3.

* binary vocabulary: W x 32 B; W/4 uniformly random "centres" plus 3 children per centre at
  Hamming distance ~ Binomial(256, 0.12) from it; ids 1..W contiguous (or strided).
* map: S signatures x F features whose words follow Zipf(s=1) over a random permutation of
  the vocabulary  =>  sum of postings ~ S*F.
* query frame: the words of one map signature ("place"), each descriptor = the word's row
  with i.i.d. bit flips (p=0.05), and a fraction replaced by random descriptors so that the
  NNDR test rejects them and the new-word path is exercised.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


def make_binary_vocabulary(n_words: int, dim_bytes: int = 32, seed: int = 1, child_p: float = 0.12) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n_centres = max(1, (n_words + 3) // 4)
    centres = rng.integers(0, 256, size=(n_centres, dim_bytes), dtype=np.uint8)
    vocab = np.empty((n_centres * 4, dim_bytes), dtype=np.uint8)
    vocab[0::4] = centres
    for c in range(1, 4):
        flips = rng.random((n_centres, dim_bytes * 8)) < child_p
        vocab[c::4] = centres ^ np.packbits(flips, axis=1)
    return np.ascontiguousarray(vocab[:n_words])


def make_float_vocabulary(n_words: int, dim: int = 64, seed: int = 1) -> np.ndarray:
    """SURF-like float vocabulary: unit L2 norm, Laplacian-distributed components, generated in chunks so that a
    million rows stay cheap."""
    rng = np.random.default_rng(seed)
    out = np.empty((n_words, dim), np.float32)
    for r0 in range(0, n_words, 1 << 16):
        r1 = min(n_words, r0 + (1 << 16))
        v = rng.laplace(0.0, 1.0, (r1 - r0, dim)).astype(np.float32)
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        out[r0:r1] = v
    return out


def flip_bits(desc: np.ndarray, p: float, rng: np.random.Generator) -> np.ndarray:
    flips = rng.random((desc.shape[0], desc.shape[1] * 8)) < p
    return desc ^ np.packbits(flips, axis=1)


@dataclass
class SynthMap:
    """Inverted index of a synthetic map in CSR form over words (ascending word id)."""
    word_ids: np.ndarray      # [n_words_with_refs] int32, ascending
    row_ptr: np.ndarray       # [n+1] int64
    sig: np.ndarray           # [nnz] int32
    cnt: np.ndarray           # [nnz] int32
    sig_ids: np.ndarray       # [S] int32 (1..S)
    ni: np.ndarray            # [S] int32 words per signature
    sig_words: np.ndarray     # [S, F] int32 word id of every feature (the forward index)

    @property
    def nnz(self) -> int:
        return int(self.row_ptr[-1])


def make_map(word_ids: np.ndarray, n_signatures: int, feats_per_sig: int, seed: int = 2, zipf_s: float = 1.0) -> SynthMap:
    rng = np.random.default_rng(seed)
    W = len(word_ids)
    ranks = np.arange(1, W + 1, dtype=np.float64)
    p = 1.0 / ranks ** zipf_s
    cdf = np.cumsum(p / p.sum())
    perm = rng.permutation(W)
    u = rng.random(n_signatures * feats_per_sig)
    r = np.minimum(np.searchsorted(cdf, u), W - 1)
    words = np.asarray(word_ids, dtype=np.int32)[perm[r]].reshape(n_signatures, feats_per_sig)
    sig_ids = np.arange(1, n_signatures + 1, dtype=np.int32)
    flat_w = words.reshape(-1).astype(np.int64)
    flat_s = np.repeat(sig_ids.astype(np.int64), feats_per_sig)
    key = flat_w * (int(n_signatures) + 2) + flat_s
    uk, cnt = np.unique(key, return_counts=True)
    pw = (uk // (int(n_signatures) + 2)).astype(np.int32)
    ps = (uk % (int(n_signatures) + 2)).astype(np.int32)
    uw, start = np.unique(pw, return_index=True)
    row_ptr = np.concatenate([start, [len(pw)]]).astype(np.int64)
    ni = np.full(n_signatures, feats_per_sig, dtype=np.int32)
    return SynthMap(uw.astype(np.int32), row_ptr, ps, cnt.astype(np.int32), sig_ids, ni, words)


def make_query_frames(vocab: np.ndarray, word_ids: np.ndarray, smap: SynthMap, n_frames: int, feats_per_frame: int,
                      seed: int = 3, flip_p: float = 0.05, new_frac: float = 0.2):
    """Frames revisiting random places of the map. Returns (desc [n_frames*F, D] uint8, place ids [n_frames])."""
    rng = np.random.default_rng(seed)
    id2row = np.full(int(np.max(word_ids)) + 1, -1, dtype=np.int64)
    id2row[np.asarray(word_ids)] = np.arange(len(word_ids))
    places = rng.integers(0, len(smap.sig_ids), size=n_frames)
    out = np.empty((n_frames * feats_per_frame, vocab.shape[1]), dtype=np.uint8)
    F = smap.sig_words.shape[1]
    for f in range(n_frames):
        sel = rng.permutation(F)[:feats_per_frame] if feats_per_frame <= F else rng.integers(0, F, feats_per_frame)
        rows = id2row[smap.sig_words[places[f], sel]]
        d = flip_bits(vocab[rows], flip_p, rng)
        n_new = int(round(new_frac * feats_per_frame))
        if n_new:
            idx = rng.permutation(feats_per_frame)[:n_new]
            d[idx] = rng.integers(0, 256, size=(n_new, vocab.shape[1]), dtype=np.uint8)
            # make a few of the random ones near-duplicates of each other so the intra-frame
            # new-word comparison (Kp/NewWordsComparedTogether) has work to do
            for k in range(0, n_new - 1, 4):
                d[idx[k + 1]] = flip_bits(d[idx[k]:idx[k] + 1], 0.03, rng)[0]
        out[f * feats_per_frame:(f + 1) * feats_per_frame] = d
    return out, smap.sig_ids[places]


# ------------------------------------------------------------------------------ geometry ----------
CAMERA_K4 = (525.0, 525.0, 320.0, 240.0)  # fx, fy, cx, cy of the 640x480 RGB-D stream


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    theta = float(np.linalg.norm(rvec))
    if theta < 1e-12:
        return np.eye(3)
    k = rvec / theta
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.cos(theta) * np.eye(3) + (1 - np.cos(theta)) * np.outer(k, k) + np.sin(theta) * Kx


def sparse_flip_mask(shape, rng: np.random.Generator, and_terms: int = 5) -> np.ndarray:
    """Random byte mask whose bits are set with probability 2^-and_terms (cheap bulk bit flips)."""
    m = np.frombuffer(rng.bytes(int(np.prod(shape))), dtype=np.uint8).reshape(shape).copy()
    for _ in range(and_terms - 1):
        m &= np.frombuffer(rng.bytes(int(np.prod(shape))), dtype=np.uint8).reshape(shape)
    return m


@dataclass
class SynthStore:
    """Per-signature data kept for verification: descriptors and 3-D points of every feature."""
    desc: np.ndarray  # [S, F, D] uint8
    xyz: np.ndarray   # [S, F, 3] float32 (NaN = no depth)


def make_signature_store(vocab: np.ndarray, word_ids: np.ndarray, smap: SynthMap, seed: int = 4, nan_frac: float = 0.02) -> SynthStore:
    rng = np.random.default_rng(seed)
    id2row = np.full(int(np.max(word_ids)) + 1, -1, dtype=np.int64)
    id2row[np.asarray(word_ids)] = np.arange(len(word_ids))
    S, F = smap.sig_words.shape
    desc = vocab[id2row[smap.sig_words.reshape(-1)]].reshape(S, F, vocab.shape[1])
    desc ^= sparse_flip_mask(desc.shape, rng, 5)  # ~3 % of the bits differ from the word's centre
    xyz = np.empty((S, F, 3), np.float32)
    z = rng.uniform(0.5, 5.0, (S, F)).astype(np.float32)
    xyz[..., 0] = rng.uniform(-0.55, 0.55, (S, F)).astype(np.float32) * z   # inside the 640x480 frustum at f=525
    xyz[..., 1] = rng.uniform(-0.40, 0.40, (S, F)).astype(np.float32) * z
    xyz[..., 2] = z
    xyz[rng.random((S, F)) < nan_frac] = np.nan
    return SynthStore(np.ascontiguousarray(desc), xyz)


def make_query_frames_geo(store: SynthStore, smap: SynthMap, n_frames: int, seed: int = 3, flip_terms: int = 4, pixel_noise: float = 0.5,
                          outlier_frac: float = 0.2, K4=CAMERA_K4):
    """Frames revisiting random places: the place's features re-observed from a nearby pose.
    Returns desc [n_frames*F, D], uv [n_frames*F, 2], place ids [n_frames], poses (rvec, tvec) [n_frames, 6]."""
    rng = np.random.default_rng(seed)
    S, F, D = store.desc.shape
    places = rng.integers(0, S, size=n_frames)
    desc = np.empty((n_frames, F, D), np.uint8)
    uv = np.empty((n_frames, F, 2), np.float32)
    poses = np.zeros((n_frames, 6))
    for f in range(n_frames):
        p = places[f]
        perm = rng.permutation(F)
        d = store.desc[p][perm] ^ sparse_flip_mask((F, D), rng, flip_terms)  # ~6 % bit flips
        X = np.nan_to_num(store.xyz[p][perm].astype(np.float64), nan=1.0)
        rv = rng.normal(0, 0.06, 3)
        tv = rng.normal(0, 0.1, 3)
        Xc = X @ rodrigues(rv).T + tv
        u = K4[0] * Xc[:, 0] / Xc[:, 2] + K4[2] + rng.normal(0, pixel_noise, F)
        v = K4[1] * Xc[:, 1] / Xc[:, 2] + K4[3] + rng.normal(0, pixel_noise, F)
        n_out = int(round(outlier_frac * F))
        idx = rng.permutation(F)[:n_out]
        d[idx] = np.frombuffer(rng.bytes(n_out * D), dtype=np.uint8).reshape(n_out, D)
        u[idx] = rng.uniform(0, 640, n_out)
        v[idx] = rng.uniform(0, 480, n_out)
        desc[f] = d
        uv[f, :, 0] = u
        uv[f, :, 1] = v
        poses[f, :3] = rv
        poses[f, 3:] = tv
    return desc.reshape(n_frames * F, D), uv.reshape(n_frames * F, 2), smap.sig_ids[places], poses


# ------------------------------------------------------------------------------ images ------------
def make_image(height: int = 480, width: int = 640, seed: int = 5, n_rects: int = 1500, bgr: bool = False) -> np.ndarray:
    """Textured synthetic frame: three octaves of smooth value noise plus random
    high-contrast rectangles so that FAST fires a few thousand times."""
    rng = np.random.default_rng(seed)
    img = np.zeros((height, width), np.float32)
    for octv, amp in ((8, 60.0), (16, 40.0), (32, 30.0)):
        small = rng.random((height // octv + 2, width // octv + 2)).astype(np.float32)
        up = np.kron(small, np.ones((octv, octv), np.float32))[:height + octv, :width + octv]
        # cheap smoothing of the blocky upsampling
        k = octv // 2
        up = (up[:-octv, :-octv] + up[k:k - octv, :-octv] + up[:-octv, k:k - octv] + up[k:k - octv, k:k - octv]) * 0.25
        img += amp * up[:height, :width]
    img = np.clip(img, 0, 255).astype(np.uint8)
    xs = rng.integers(0, width - 30, n_rects)
    ys = rng.integers(0, height - 30, n_rects)
    ws = rng.integers(4, 30, n_rects)
    hs = rng.integers(4, 30, n_rects)
    vals = rng.integers(0, 256, n_rects)
    for x, y, w_, h_, v in zip(xs, ys, ws, hs, vals):
        img[y:y + h_, x:x + w_] = v
    if not bgr:
        return img
    out = np.stack([img, np.roll(img, 3, axis=1), np.roll(img, 5, axis=0)], axis=2)
    out ^= (rng.integers(0, 8, out.shape, dtype=np.uint8))
    return np.ascontiguousarray(out)


def make_depth(height: int = 480, width: int = 640, seed: int = 6, zero_frac: float = 0.02, as_float: bool = False) -> np.ndarray:
    """Depth of a slanted plane at 1-4 m with 5 mm noise and a few invalid pixels / blocks (CV_16UC1 mm or CV_32FC1 m)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    z = 1.0 + 3.0 * (0.6 * xx / width + 0.4 * yy / height) + rng.normal(0, 0.005, (height, width)).astype(np.float32)
    mm = np.clip(z * 1000.0, 1, 65000).astype(np.uint16)
    mm[rng.random((height, width)) < zero_frac] = 0
    for _ in range(6):
        x, y = rng.integers(0, width - 80), rng.integers(0, height - 60)
        mm[y:y + rng.integers(20, 60), x:x + rng.integers(20, 80)] = 0
    if as_float:
        f = mm.astype(np.float32) * np.float32(0.001)
        f[mm == 0] = np.nan if seed % 2 else 0.0
        return f
    return mm


# ------------------------------------------------------------------------------ image-level world --
def render_view(image: np.ndarray, depth: np.ndarray, dx: int, dy: int, noise_sigma: float, rng: np.random.Generator):
    """A revisit of a place: the frame shifted by (dx, dy) pixels (edge-replicated) with sensor noise."""
    h, w = depth.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    img = image[ys][:, xs]
    dep = depth[ys][:, xs]
    if noise_sigma > 0:
        noise = rng.normal(0, noise_sigma, img.shape)
        img = np.clip(img.astype(np.float32) + noise, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(img), np.ascontiguousarray(dep)


@dataclass
class PlaceWorld:
    """A map built from P textured places seen S times (BASELINE configs[1] at image level):
    the vocabulary is made of the places' own ORB descriptors, every signature is a view of a place."""
    images: np.ndarray        # [P, H, W, 3] uint8 BGR
    depths: np.ndarray        # [P, H, W] uint16 mm
    vocab: np.ndarray         # [W, 32] uint8
    word_ids: np.ndarray      # [W] int32 (1..W)
    smap: SynthMap            # inverted index of the S signatures
    store: SynthStore         # per-signature descriptors and 3-D points
    sig_place: np.ndarray     # [S] place of every signature


def make_place_world(orb_fn, n_places: int = 50, n_words: int = 49152, n_signatures: int = 10000, feats: int = 1000,
                     height: int = 480, width: int = 640, seed: int = 7) -> PlaceWorld:
    """orb_fn(image_bgr, depth) -> (keypoints [n,6], descriptors [n,32], xyz [n,3]) is supplied by the caller
    (the benchmark uses the OpenCV-based checker for this untimed set-up)."""
    rng = np.random.default_rng(seed)
    images = np.stack([make_image(height, width, 100 + p, bgr=True) for p in range(n_places)])
    depths = np.stack([make_depth(height, width, 200 + p, zero_frac=0.01) for p in range(n_places)])
    desc = np.zeros((n_places, feats, 32), np.uint8)
    xyz = np.full((n_places, feats, 3), np.nan, np.float32)
    cnt = np.zeros(n_places, np.int64)
    for p in range(n_places):
        kp, d, x = orb_fn(images[p], depths[p])
        n = min(len(d), feats)
        desc[p, :n], xyz[p, :n], cnt[p] = d[:n], x[:n], n
    # vocabulary: the places' descriptors, in place order
    flat_place = np.repeat(np.arange(n_places), feats)
    flat_idx = np.tile(np.arange(feats), n_places)
    valid = flat_idx < cnt[flat_place]
    all_desc = desc.reshape(-1, 32)[valid]
    W = min(n_words, len(all_desc))
    vocab = np.ascontiguousarray(all_desc[:W])
    word_ids = np.arange(1, W + 1, dtype=np.int32)
    word_of = np.zeros((n_places, feats), np.int32)  # 0 = the descriptor is not a vocabulary word
    word_of.reshape(-1)[np.nonzero(valid)[0][:W]] = word_ids
    # signatures: S views of the places, round-robin
    sig_place = (np.arange(n_signatures) % n_places).astype(np.int64)
    sig_ids = np.arange(1, n_signatures + 1, dtype=np.int32)
    sig_words = word_of[sig_place]                                  # [S, feats]
    sdesc = desc[sig_place] ^ sparse_flip_mask((n_signatures, feats, 32), rng, 6)   # ~1.6 % bit noise per view
    sxyz = xyz[sig_place] + rng.normal(0, 0.002, (n_signatures, feats, 3)).astype(np.float32)
    ni = cnt[sig_place].astype(np.int32)
    # inverted index (word -> signatures), every occurrence counts once per feature
    fw = sig_words.reshape(-1).astype(np.int64)
    fs = np.repeat(sig_ids.astype(np.int64), feats)
    keep = fw > 0
    key = fw[keep] * (int(n_signatures) + 2) + fs[keep]
    uk, c = np.unique(key, return_counts=True)
    pw = (uk // (int(n_signatures) + 2)).astype(np.int32)
    ps = (uk % (int(n_signatures) + 2)).astype(np.int32)
    uw, start = np.unique(pw, return_index=True)
    row_ptr = np.concatenate([start, [len(pw)]]).astype(np.int64)
    smap = SynthMap(uw.astype(np.int32), row_ptr, ps, c.astype(np.int32), sig_ids, ni, sig_words)
    # rows beyond a place's keypoint count carry no data
    for p in range(n_places):
        if cnt[p] < feats:
            sel = sig_place == p
            sxyz[sel, cnt[p]:] = np.nan
    return PlaceWorld(images, depths, vocab, word_ids, smap, SynthStore(np.ascontiguousarray(sdesc), sxyz.astype(np.float32)), sig_place)


def warp_view(image: np.ndarray, depth: np.ndarray, angle_deg: float, scale: float, dx: float, dy: float, noise_sigma: float,
              rng: np.random.Generator):
    """A revisit from a different viewpoint ("rotated / homography views"): the frame rotated about its centre,
    scaled and shifted (bilinear for the image, nearest for the depth, which is divided by the scale: closer = larger)."""
    import cv2  # input synthesis only

    h, w = depth.shape
    M = cv2.getRotationMatrix2D((w * 0.5, h * 0.5), angle_deg, scale)
    M[0, 2] += dx
    M[1, 2] += dy
    img = cv2.warpAffine(image, M, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    dep = cv2.warpAffine(depth, M, (w, h), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    dep = np.where(dep > 0, np.clip(dep.astype(np.float32) / np.float32(scale), 1, 65000), 0).astype(np.uint16)
    if noise_sigma > 0:
        img = np.clip(img.astype(np.float32) + rng.normal(0, noise_sigma, img.shape), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(img), np.ascontiguousarray(dep)


def make_view_frames(world: PlaceWorld, n_frames: int, seed: int = 3, max_shift: int = 6, noise_sigma: float = 2.0, mode: str = "revisit",
                     new_place_frac: float = 0.3, warp_frac: float = 0.4):
    """Query frames.  mode "revisit": noisy, slightly shifted views of random places (every query is a true loop closure).
    mode "mixed" (the harder workload SURVEY.md §8(d) specifies): `new_place_frac` of the frames show places that are NOT in the map
    (no loop closure: the likelihood is flat, verification runs all its RANSAC iterations and rejects), `warp_frac` are rotated
    (+-12 deg) / scaled (0.85-1.15) / shifted views, the rest plain shifted revisits.
    Returns images [n,H,W,3], depths [n,H,W], places [n] (-1 = never-seen place)."""
    rng = np.random.default_rng(seed)
    P = len(world.images)
    places = rng.integers(0, P, n_frames)
    imgs = np.empty((n_frames,) + world.images.shape[1:], np.uint8)
    deps = np.empty((n_frames,) + world.depths.shape[1:], np.uint16)
    h, w = world.depths.shape[1:]
    kind = np.zeros(n_frames, np.int64)
    if mode == "mixed":
        u = rng.random(n_frames)
        kind = np.where(u < new_place_frac, 2, np.where(u < new_place_frac + warp_frac, 1, 0))
    elif mode != "revisit":
        raise ValueError(mode)
    for f in range(n_frames):
        dx, dy = rng.integers(-max_shift, max_shift + 1, 2)
        if kind[f] == 2:
            places[f] = -1
            img = make_image(h, w, 100000 + seed * 1000 + f, bgr=True)
            dep = make_depth(h, w, 200000 + seed * 1000 + f, zero_frac=0.01)
            imgs[f], deps[f] = render_view(img, dep, int(dx), int(dy), noise_sigma, rng)
        elif kind[f] == 1:
            imgs[f], deps[f] = warp_view(world.images[places[f]], world.depths[places[f]], float(rng.uniform(-12, 12)), float(rng.uniform(0.85, 1.15)),
                                         float(rng.uniform(-20, 20)), float(rng.uniform(-20, 20)), noise_sigma, rng)
        else:
            imgs[f], deps[f] = render_view(world.images[places[f]], world.depths[places[f]], int(dx), int(dy), noise_sigma, rng)
    return imgs, deps, places
