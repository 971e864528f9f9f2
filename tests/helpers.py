"""Shared test helpers: random setups / packets, and the oracle-driven reference decode."""
import numpy as np

import lewton_b200 as L


def bits_equal(a, b):
    """Bit-identical float arrays, treating +0/-0 as equal and any-NaN == any-NaN
    (SURVEY.md section 8c parity rule)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    same = a.view(np.uint32) == b.view(np.uint32)
    zeros = (a == 0) & (b == 0)
    nans = np.isnan(a) & np.isnan(b)
    return bool(np.all(same | zeros | nans))


def mismatch_report(a, b):
    a = np.ascontiguousarray(a, np.float32).ravel()
    b = np.ascontiguousarray(b, np.float32).ravel()
    bad = np.nonzero(~((a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)) | (np.isnan(a) & np.isnan(b))))[0]
    return f"{bad.size} of {a.size} differ; first at {bad[:5]}: got {a[bad[:5]]} want {b[bad[:5]]}"


def random_floor1(rng, n2):
    mult = int(rng.integers(1, 5))
    rangebits = int(rng.integers(max(4, int(np.log2(n2)) - 2), min(15, int(np.log2(n2)) + 2) + 1))
    nposts = int(rng.integers(2, 66))
    nposts = min(nposts, 1 << rangebits)
    xs = [0, 1 << rangebits] + [int(v) for v in rng.permutation(np.arange(1, 1 << rangebits))[: nposts - 2]]
    return mult, xs


def random_floor1_y(rng, mult, nposts, wild=False):
    rng_y = [256, 128, 86, 64][mult - 1]
    y = [int(rng.integers(0, rng_y)), int(rng.integers(0, rng_y))]
    for _ in range(nposts - 2):
        r = rng.random()
        if r < 0.3:
            y.append(0)
        elif r < 0.9 or not wild:
            y.append(int(rng.integers(1, 40)))
        else:
            y.append(int(rng.integers(1, 5000)))
    return y


class RefStream:
    """Oracle-side twin of one stream: decodes packet by packet with oracle.synth_*."""

    def __init__(self, oracle, channels, bs0, bs1, modes, mappings=None, floors=None):
        self.o, self.ch, self.bs0, self.bs1 = oracle, channels, bs0, bs1
        self.modes = modes                      # [(blockflag, mapping)]
        self.mappings = mappings or [{"coupling": [], "floor_of_channel": [0] * channels}]
        self.floors = floors or []              # [(mult, xs)] -> oracle Floor1 objects
        self.ofloors = [oracle.make_floor1(m, xs) for (m, xs) in self.floors]
        self.pwr = oracle.Pwr(channels, bs1)

    def spectrum(self, mode, prev, nxt, spec):
        bf = self.modes[mode][0]
        return self.o.synth_spectrum(self.bs0, self.bs1, bf, prev, nxt, spec, self.pwr)

    def packet(self, mode, prev, nxt, residue, floors):
        """floors: per channel None | list y | ndarray dense"""
        bf, mi = self.modes[mode]
        mp = self.mappings[mi]
        fl = []
        for c, f in enumerate(floors):
            if f is None or (isinstance(f, np.ndarray) and f.dtype.kind == "f"):
                fl.append(f)
            else:
                fl.append((self.ofloors[mp["floor_of_channel"][c]], f))
        return self.o.synth_packet(self.bs0, self.bs1, bf, prev, nxt, mp["coupling"], fl, residue, self.pwr)


def make_setup(ctx, channels, bs0, bs1, modes=((0, 0), (1, 0)), mappings=None, floors=None, tables=None):
    """modes: [(blockflag, mapping)]; mappings: [{"coupling": [(m,a)..], "floor_of_channel": [..]}];
    floors: [(mult, xs)]"""
    mappings = mappings or [{"coupling": [], "floor_of_channel": [0] * channels}]
    floors = floors or [(1, [0, 128])]
    lf = [L.FloorTypeOne(m, xs) for (m, xs) in floors]
    lm = []
    for mp in mappings:
        # one submap per distinct floor, mux maps channel -> submap
        fo = mp["floor_of_channel"]
        uniq = sorted(set(fo))
        lm.append(L.Mapping(channels, [m for m, _ in mp["coupling"]], [a for _, a in mp["coupling"]],
                            mux=[uniq.index(f) for f in fo], submap_floors=uniq))
    lmodes = [L.ModeInfo(bf, mi) for bf, mi in modes]
    return L.Setup(ctx, channels, bs0, bs1, lf, lm, lmodes, tables=tables)


def mode_sequence(rng, n, p_short=0.3):
    """Random block-type sequence with consistent prev/next flags: returns (modes, prev, next)
    with mode 0 = short, 1 = long."""
    bf = (rng.random(n) >= p_short).astype(np.uint8)
    prev = np.ones(n, np.uint8)
    nxt = np.ones(n, np.uint8)
    for i in range(n):
        if bf[i]:
            prev[i] = bf[i - 1] if i else 1
            nxt[i] = bf[i + 1] if i + 1 < n else 1
    return bf, prev, nxt
