"""Oracle ZSobol sampler against the known-answer vectors of SURVEY.md Appendix B (tests/golden/sobol_kat.json) and against an
independent pure-Python transcription of z_sobol_sampler.rs kept in this file; GPU stream vs oracle, bit-exact."""
import json
import struct
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "sobol_kat.json").read_text())["rows"]
M64 = (1 << 64) - 1


def _oracle(tables):
    from oracle import oracle
    return oracle.OracleScene(tables[0], tables[1])


def bits(f):
    return struct.unpack("<I", struct.pack("<f", float(f)))[0]


# ---- independent transcription (Python ints) of renderer/src/sampler/z_sobol_sampler.rs:3-235
PERM = [[0, 1, 2, 3], [0, 1, 3, 2], [0, 2, 1, 3], [0, 2, 3, 1], [0, 3, 2, 1], [0, 3, 1, 2], [1, 0, 2, 3], [1, 0, 3, 2], [1, 2, 0, 3], [1, 2, 3, 0], [1, 3, 2, 0], [1, 3, 0, 2],
        [2, 1, 0, 3], [2, 1, 3, 0], [2, 0, 1, 3], [2, 0, 3, 1], [2, 3, 0, 1], [2, 3, 1, 0], [3, 1, 2, 0], [3, 1, 0, 2], [3, 2, 1, 0], [3, 2, 0, 1], [3, 0, 2, 1], [3, 0, 1, 2]]


def py_mix(v):
    v ^= v >> 31; v = (v * 0x7fb5d329728ea185) & M64; v ^= v >> 27; v = (v * 0x81dadef4bc2dd44d) & M64; v ^= v >> 33
    return v


def py_hash(dim, seed):
    M, R = 0xc6a4a7935bd1e995, 47
    h = (8 * M) & M64
    k = dim | (seed << 32)
    k = (k * M) & M64; k ^= k >> R; k = (k * M) & M64
    h ^= k; h = (h * M) & M64
    h ^= h >> R; h = (h * M) & M64; h ^= h >> R
    return h


def py_rev(n):
    return int(f"{n:032b}"[::-1], 2)


def py_owen(v, seed):
    m = 0xffffffff
    v = py_rev(v)
    v ^= (v * 0x3d20adea) & m; v = (v + seed) & m; v = (v * ((seed >> 16) | 1)) & m; v ^= (v * 0x05526c56) & m; v ^= (v * 0x53a22864) & m
    return py_rev(v)


def py_shift2(x):
    x &= 0xffffffff
    for s, msk in ((16, 0x0000ffff0000ffff), (8, 0x00ff00ff00ff00ff), (4, 0x0f0f0f0f0f0f0f0f), (2, 0x3333333333333333), (1, 0x5555555555555555)):
        x = (x ^ (x << s)) & msk
    return x


class PySobol:
    def __init__(self, mats, spp, w, h, seed):
        self.m, self.seed = mats, seed
        self.log2 = spp.bit_length() - 1 if spp else 0
        res = 1 if max(w, h) <= 1 else 1 << (max(w, h) - 1).bit_length()
        self.nb4 = (res.bit_length() - 1) + (self.log2 + 1) // 2
        self.dim = 0

    def start(self, px, py, i):
        self.dim = 0
        mort = (((py_shift2(py) & 0xffffffff) << 1) & 0xffffffff) | (py_shift2(px) & 0xffffffff)
        self.morton = ((mort << self.log2) & 0xffffffff) | i

    def index(self):
        s, pow2 = 0, (self.log2 & 1) == 1
        last = 1 if pow2 else 0
        i = self.nb4 - 1
        while i >= last:
            sh = 2 * i - (1 if pow2 else 0)
            d = (self.morton >> sh) & 3
            p = (py_mix((self.morton >> (sh + 2)) ^ ((0x55555555 * self.dim) & M64)) >> 24) % 24
            s |= PERM[p][d] << sh
            i -= 1
        if pow2:
            d = self.morton & (i & M64)
            s |= d ^ (py_mix((self.morton >> 1) ^ ((0x55555555 * self.dim) & M64)) & 1)
        return s

    def sample(self, a, dim, scr):
        v, i = 0, dim * 52
        while a:
            if a & 1:
                v ^= int(self.m[i])
            a >>= 1; i += 1
        v = py_owen(v, scr)
        return min(np.float32(v) * np.float32(2.0 ** -32), np.float32(struct.unpack("<f", struct.pack("<I", 0x3f7fffff))[0]))

    def get_1d(self):
        a = self.index(); self.dim += 1
        return self.sample(a, 0, py_hash(self.dim, self.seed) & 0xffffffff)

    def get_2d(self):
        a = self.index(); self.dim += 2
        h = py_hash(self.dim, self.seed)
        return self.sample(a, 0, h & 0xffffffff), self.sample(a, 1, h >> 32)


def sobol_matrices(tables):
    return np.frombuffer(tables[0][8:8 + 104 * 4], dtype="<u4")


@pytest.mark.parametrize("row", GOLD, ids=lambda r: f"{r['spp']}_{r['p'][0]}_{r['p'][1]}_{r['i']}")
def test_oracle_matches_known_answers(tables, row):
    o = _oracle(tables)
    vals, idx, morton = o.sobol_probe(row["spp"], row["w"], row["h"], 0, row["p"][0], row["p"][1], row["i"])
    assert morton == int(row["morton"], 16)
    assert [int(x) for x in idx] == row["idx"]
    assert [bits(v) for v in vals] == [int(b, 16) for b in row["bits"]]


def test_python_transcription_matches_known_answers(tables):
    m = sobol_matrices(tables)
    for row in GOLD:
        s = PySobol(m, row["spp"], row["w"], row["h"], 0)
        s.start(row["p"][0], row["p"][1], row["i"])
        a = s.get_1d(); b, c = s.get_2d(); d = s.get_1d()
        assert [bits(v) for v in (a, b, c, d)] == [int(x, 16) for x in row["bits"]]


def test_oracle_matches_transcription_on_long_streams(tables):
    m = sobol_matrices(tables)
    o = _oracle(tables)
    rng = np.random.default_rng(0)
    kinds = [1, 2] + [1, 2, 1, 1, 2, 1] * 6
    for spp, w, h in ((512, 200, 150), (64, 33, 7), (1024, 1920, 1080), (4096, 3840, 2160), (1, 4, 4), (6, 10, 10)):
        for _ in range(6):
            px, py, i, seed = int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(0, spp)), int(rng.integers(0, 2 ** 32))
            got = o.sampler_stream("sobol", spp, w, h, seed, px, py, i, kinds)
            s = PySobol(m, spp, w, h, seed); s.start(px, py, i)
            exp = []
            for k in kinds:
                exp += [s.get_1d()] if k == 1 else list(s.get_2d())
            assert got.view(np.uint32).tolist() == np.array(exp, dtype=np.float32).view(np.uint32).tolist()


def test_duplicate_pairs_quirk_for_odd_log2_spp(tables):
    """Appendix A q15-i: with odd log2(spp) samples 2k and 2k+1 share every Sobol index (the reference ANDs with 0)."""
    o = _oracle(tables)
    kinds = [1, 2, 1, 2]
    a = o.sampler_stream("sobol", 512, 200, 150, 0, 11, 22, 40, kinds)
    b = o.sampler_stream("sobol", 512, 200, 150, 0, 11, 22, 41, kinds)
    assert np.array_equal(a, b)
    c = o.sampler_stream("sobol", 1024, 200, 150, 0, 11, 22, 40, kinds)
    d = o.sampler_stream("sobol", 1024, 200, 150, 0, 11, 22, 41, kinds)
    assert not np.array_equal(c, d)


@pytest.mark.gpu
@pytest.mark.parametrize("sampler", ["sobol", "random"])
def test_gpu_stream_bit_exact(tables, sampler):
    import ctypes as C
    from toy_cpu_pathtracing_b200 import capi
    ctx = capi.Context(0)
    o = _oracle(tables)
    rng = np.random.default_rng(3)
    kinds = np.array([1, 2] + [1, 2, 1, 1, 2, 1] * 16, dtype=np.int32)
    total = int(sum(1 if k == 1 else 2 for k in kinds))
    cases = [(r["spp"], r["w"], r["h"], r["p"][0], r["p"][1], r["i"], 0) for r in GOLD]
    for spp, w, h in ((512, 200, 150), (1024, 1920, 1080), (4096, 3840, 2160), (2048, 200, 150), (1, 3, 3), (6, 10, 10)):
        for _ in range(8):
            cases.append((spp, w, h, int(rng.integers(0, w)), int(rng.integers(0, h)), int(rng.integers(0, spp)), int(rng.integers(0, 2 ** 32))))
    for spp, w, h, px, py, i, seed in cases:
        out = np.zeros(total, dtype=np.float32)
        n = ctx.check(ctx.lib.tcpt_sampler_stream(ctx.handle, capi.SAMPLERS[sampler], spp, w, h, seed, px, py, i, capi.as_ptr(kinds, C.c_int32), len(kinds), capi.as_ptr(out, C.c_float)))
        assert n == total
        exp = o.sampler_stream(sampler, spp, w, h, seed, px, py, i, kinds)
        assert out.view(np.uint32).tolist() == exp.view(np.uint32).tolist(), (sampler, spp, w, h, px, py, i, seed)


@pytest.mark.parametrize("spp", [16, 64, 256])
def test_samples_of_a_pixel_form_a_base2_net(tables, spp):
    """Ground truth that does not depend on any restatement: the samples of one pixel are an aligned block of 2^m consecutive Sobol
    indices (the Morton prefix selects the block, the digit permutations only reorder it), and Owen scrambling keeps the net property,
    so the 2-D points of ANY get_2d call form a (0, m, 2)-net: every elementary interval of area 2^-m holds exactly one of them.  A
    wrong generator matrix, a wrong bit order, a broken scramble or a permutation that leaves the block would all break this.
    (Even log2(spp) only: for odd ones the reference draws every sample twice, z_sobol_sampler.rs:131-155, see the quirk test above.)"""
    o = _oracle(tables)
    m = spp.bit_length() - 1
    assert m % 2 == 0
    kinds = [1, 2, 1, 2, 2]                      # wavelength, pixel sample, lobe choice, BSDF sample, light sample
    for px, py, seed in ((0, 0, 0), (13, 7, 0), (199, 149, 5)):
        pts = np.array([o.sampler_stream("sobol", spp, 200, 150, seed, px, py, i, kinds) for i in range(spp)], dtype=np.float64)
        for cols in ((1, 2), (4, 5), (6, 7)):     # the three get_2d calls
            xy = pts[:, cols]
            assert xy.min() >= 0.0 and xy.max() < 1.0
            for a in range(m + 1):
                cells = np.floor(xy[:, 0] * (1 << a)).astype(np.int64) * (1 << (m - a)) + np.floor(xy[:, 1] * (1 << (m - a))).astype(np.int64)
                assert len(np.unique(cells)) == spp, (px, py, cols, a)
        # and every 1-D draw is stratified into spp equal intervals
        for c in (0, 3):
            assert len(np.unique(np.floor(pts[:, c] * spp).astype(np.int64))) == spp
