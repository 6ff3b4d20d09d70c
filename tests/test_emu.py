"""CPU tests of the *device source* (csrc/mont52.cuh, csrc/paillier_items.cuh) compiled for the host and run
under a lockstep lane emulator (tests/emu/emu_driver.cpp), checked bit-exactly against the Python-int oracle.
These catch logic errors in the kernels before any GPU time is spent; the -m gpu tests repeat the checks on
the real CUDA build through the C ABI."""
import ctypes
import random

import numpy as np
import pytest

import paillier_oracle as O
from emu_util import P, P64, PD, from_entry, from_words, load_emu, mont_consts, shape_id, to_entry, to_words

U64 = ctypes.c_uint64


@pytest.fixture(scope="module")
def emu():
    return load_emu()


SHAPES = [(20, 1, 1024), (20, 2, 2048), (20, 4, 4096), (15, 4, 3072), (15, 8, 6144), (20, 8, 8192), (7, 4, 1408)]


@pytest.mark.parametrize("L,TPI,bits", SHAPES)
def test_modmul(emu, L, TPI, bits):
    rng = random.Random(bits + TPI)
    nw = bits // 32
    for top in (bits, bits - 5):
        N = rng.getrandbits(top) | 1 | (1 << (top - 1))
        mc = mont_consts(N, L, TPI)
        a = [0, 1, N - 1, N - 1] + [rng.randrange(N) for _ in range(6)]
        b = [5, 1, N - 1, 1] + [rng.randrange(N) for _ in range(6)]
        aw, bw = to_words(a, nw), to_words(b, nw)
        out = np.zeros_like(aw)
        assert emu.emu_modmul(shape_id(L, TPI), P(aw), P(bw), P(out), nw, len(a), PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"])) == 0
        assert from_words(out) == [x * y % N for x, y in zip(a, b)]


@pytest.mark.parametrize("L,TPI,bits", [(20, 2, 2048), (20, 4, 4096), (15, 8, 6144)])
def test_modmul_one_product(emu, L, TPI, bits):
    """item_modmul1: a * b * R^-1 with b as words or as a constant entry -- the broadcast add (b = other * R), a level of
    the add tree (the R^-1 stay in, the root repays them with R^width) and the pass-through by the Montgomery one."""
    rng = random.Random(bits + 7 * TPI)
    nw = bits // 32
    N = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    mc = mont_consts(N, L, TPI)
    R, Rinv = mc["R"], pow(mc["R"], -1, N)
    a = [0, 1, N - 1] + [rng.randrange(N) for _ in range(5)]
    b = [N - 1, 5, N - 1] + [rng.randrange(N) for _ in range(5)]
    aw, bw = to_words(a, nw), to_words(b, nw)
    out = np.zeros_like(aw)
    assert emu.emu_modmul1(shape_id(L, TPI), P(aw), P(bw), None, P(out), nw, len(a), PD(mc["n"]), U64(mc["n0inv"])) == 0
    assert from_words(out) == [x * y * Rinv % N for x, y in zip(a, b)]
    # constant entry: b = R^2 brings a into the Montgomery domain; b = R mod N is the identity
    assert emu.emu_modmul1(shape_id(L, TPI), P(aw), None, PD(mc["r2"]), P(out), nw, len(a), PD(mc["n"]), U64(mc["n0inv"])) == 0
    assert from_words(out) == [x * R % N for x in a]
    assert emu.emu_modmul1(shape_id(L, TPI), P(aw), None, PD(mc["oneM"]), P(out), nw, len(a), PD(mc["n"]), U64(mc["n0inv"])) == 0
    assert from_words(out) == a
    # a whole tree in Python on top of the emulated product: 5 leaves, root repaid with R^5
    leaves = [rng.randrange(N) for _ in range(5)]

    def prod1(x, y=None, entry=None):
        o = np.zeros((1, nw), dtype=np.uint32)
        assert emu.emu_modmul1(shape_id(L, TPI), P(to_words([x], nw)), P(to_words([y], nw)) if y is not None else None,
                               PD(entry) if entry is not None else None, P(o), nw, 1, PD(mc["n"]), U64(mc["n0inv"])) == 0
        return from_words(o)[0]
    from emu_util import to_entry
    l1 = [prod1(leaves[0], leaves[2]), prod1(leaves[1], leaves[3]), prod1(leaves[4], entry=mc["oneM"])]
    l2 = [prod1(l1[0], l1[1]), prod1(l1[2], entry=mc["oneM"])]
    root = prod1(l2[0], l2[1])
    want = 1
    for v in leaves:
        want = want * v % N
    assert prod1(root, entry=to_entry(pow(R, 5, N), L, TPI)) == want


@pytest.mark.parametrize("L,TPI,bits", [(20, 1, 1024), (20, 2, 2048), (15, 4, 3072)])
def test_npair_scale_rows(emu, L, TPI, bits):
    """NPairScaleCtl: c^(2^delta) by delta squarings; a lane group whose delta is below the batch maximum keeps the value
    it had after its own delta (exponent alignment, ipcl_python.py:551-560)."""
    rng = random.Random(bits * 5 + 1)
    nw = bits // 32
    n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    n2 = n * n
    nc = npair_consts(n, L, TPI, nw)
    base = [rng.randrange(n2) for _ in range(4)] + [1, n2 - 1, n + 1]
    delta = np.array([9, 0, 3, 9, 5, 1, 8], dtype=np.int32)
    cw = to_words(base, 2 * nw)
    out = np.zeros_like(cw)
    rc = emu.emu_scale_npair(shape_id(L, TPI), P(cw), nw, delta.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), 9, P(out),
                             len(base), PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"]))
    assert rc == 0
    assert from_words(out) == [pow(b, 1 << int(d), n2) for b, d in zip(base, delta)]


@pytest.mark.parametrize("L,TPI,bits,win,ebits", [(20, 1, 1024, 5, 300), (20, 2, 2048, 5, 1024), (20, 4, 4096, 3, 53),
                                                   (20, 4, 4096, 1, 7), (20, 2, 2048, 3, 1), (15, 4, 3072, 5, 100)])
def test_powm_per_item_exponent(emu, L, TPI, bits, win, ebits):
    rng = random.Random(bits * 7 + win)
    nw = bits // 32
    ew = (ebits + 31) // 32
    N = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    mc = mont_consts(N, L, TPI)
    base = [rng.randrange(N) for _ in range(3)] + [0, 1]
    exps = [rng.getrandbits(ebits) for _ in range(3)] + [3, 0]
    exps[0] |= 1 << (ebits - 1)
    bw, ewa = to_words(base, nw), to_words(exps, ew)
    out = np.zeros_like(bw)
    rc = emu.emu_powm(shape_id(L, TPI), win, P(bw), nw, None, P(ewa), ew, ew, ebits, P(out), nw, len(base),
                      PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
    assert rc == 0
    assert from_words(out) == [pow(b, e, N) for b, e in zip(base, exps)]


def run_program(prog, x, N):
    """Reference interpreter of a sliding-window program (include/phe_b200.h: phe_host_powm_program)."""
    if prog[0] == 0xFFFF:
        return 1 % N
    acc = pow(x, 2 * prog[0] + 1, N)
    for op in prog[1:]:
        acc = pow(acc, 1 << (op >> 8), N)
        if op & 0xFF != 0xFF:
            acc = acc * pow(x, 2 * (op & 0xFF) + 1, N) % N
    return acc


@pytest.mark.parametrize("L,TPI,bits,ebits", [(20, 1, 1024, 512), (20, 2, 2048, 1024), (15, 4, 3072, 200), (20, 1, 512, 9)])
def test_powm_shared_exponent_program(emu, L, TPI, bits, ebits):
    from pailliercryptolib_python_b200 import capi
    rng = random.Random(bits + ebits)
    nw = bits // 32
    N = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    mc = mont_consts(N, L, TPI)
    base = [rng.randrange(N) for _ in range(3)] + [0, 1, N - 1]
    bw = to_words(base, nw)
    for e in (rng.getrandbits(ebits) | (1 << (ebits - 1)), (1 << ebits) - 1, 1 << (ebits - 1), 1, 0, 0b1000001000000):
        prog = capi.host_powm_program(e, (ebits + 31) // 32 + 1)
        assert [run_program(prog, b, N) for b in base] == [pow(b, e, N) for b in base]
        assert sum(op >> 8 for op in prog[1:]) <= max(e.bit_length() - 1, 0)
        pa = np.array(prog, dtype=np.uint32)
        out = np.zeros_like(bw)
        rc = emu.emu_powm_prog(shape_id(L, TPI), P(bw), nw, None, P(pa), len(prog) - 1, P(out), nw, len(base),
                               PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
        assert rc == 0
        assert from_words(out) == [pow(b, e, N) for b in base]


def _dec_consts(sk, L, TPI):
    """Decrypt-tail constant block, computed with Python ints."""
    R = 1 << (52 * L * TPI)
    p, q, n = sk.p, sk.q, sk.pk.n
    ents = [p, q, n, sk.hp * R % p, sk.hq * R % q, sk.pinv * R % q, p * R % n, 1]
    cst = np.concatenate([to_entry(v, L, TPI) for v in ents])
    n0 = np.array([(-pow(m, -1, 1 << 52)) % (1 << 52) for m in (p, q, n)], dtype=np.uint64)
    return cst, n0


@pytest.mark.parametrize("bits,L,TPI", [(1024, 20, 1), (2048, 20, 2), (3072, 15, 4)])
def test_decrypt_pipeline(emu, bits, L, TPI):
    """dec_prep -> powm(shared exponent x-1) -> dec_tail == oracle decrypt_crt."""
    if bits == 2048:
        pk, sk = O.bench_keypair()
    else:
        pk, sk = O.seeded_keypair(bits, 11)
    rng = random.Random(bits)
    ms = [0, 1, pk.n - 1] + [rng.randrange(pk.n) for _ in range(2)]
    cs = [O.encrypt(pk, m, rng.getrandbits(bits // 2)) for m in ms] + [0, sk.p, pk.n]   # and three non-units
    cw = to_words(cs, bits // 16)
    hw = bits // 32
    us = []
    for x in (sk.p, sk.q):
        X2 = x * x
        mc = mont_consts(X2, L, TPI)
        k2 = to_entry((1 << (32 * hw)) * mc["R"] * mc["R"] % X2, L, TPI)
        KP = len(mc["n"])
        ent = np.zeros((len(cs), KP), dtype=np.float64)
        assert emu.emu_dec_prep(shape_id(L, TPI), P(cw), hw, PD(ent), len(cs), PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(k2)) == 0
        for i, c in enumerate(cs):
            assert from_entry(ent[i], L, TPI) % X2 == c * mc["R"] % X2
        e = to_words([x - 1], hw)
        out = np.zeros((len(cs), hw), dtype=np.uint32)
        rc = emu.emu_powm(shape_id(L, TPI), 5, None, 0, PD(ent), P(e), hw, 0, (x - 1).bit_length(), P(out), hw, len(cs),
                          PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
        assert rc == 0
        assert from_words(out) == [pow(c % X2, x - 1, X2) for c in cs]
        from pailliercryptolib_python_b200 import capi
        pa = np.array(capi.host_powm_program(x - 1, hw), dtype=np.uint32)
        out2 = np.zeros_like(out)
        rc = emu.emu_powm_prog(shape_id(L, TPI), None, 0, PD(ent), P(pa), len(pa) - 1, P(out2), hw, len(cs),
                               PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
        assert rc == 0 and np.array_equal(out, out2)
        us.append(out)
    cst, n0 = _dec_consts(sk, L, TPI)
    mo = np.zeros((len(cs), hw), dtype=np.uint32)
    assert emu.emu_dec_tail(shape_id(L, TPI), P(us[0]), P(us[1]), hw, P(mo), hw, len(cs), PD(cst), P64(n0)) == 0
    assert from_words(mo) == O.decrypt_batch(sk, cs) and from_words(mo)[:len(ms)] == ms


def _comb_table(pk, L, TPI, WB=8):
    N = pk.nsquare
    R = 1 << (52 * L * TPI)
    nwin = (pk.randbits + WB - 1) // WB
    rows = []
    for j in range(nwin):
        base = pow(pk.hs, 1 << (WB * j), N)
        v = 1
        for d in range(1 << WB):
            rows.append(to_entry(v * R % N, L, TPI))
            v = v * base % N
    return np.concatenate(rows), nwin


@pytest.mark.parametrize("bits,L,TPI,WB", [(1024, 20, 2, 8), (1024, 20, 2, 5), (1024, 20, 2, 11)])
def test_encrypt_comb(emu, bits, L, TPI, WB):
    pk, sk = O.seeded_keypair(bits, 3)
    N = pk.nsquare
    mc = mont_consts(N, L, TPI)
    nR = to_entry(pk.n * mc["R"] % N, L, TPI)
    comb, nwin = _comb_table(pk, L, TPI, WB)
    rng = random.Random(9)
    ms = [0, 1, pk.n - 1, rng.randrange(pk.n), rng.getrandbits(53)]
    rs = [0, 1, (1 << pk.randbits) - 1, rng.getrandbits(pk.randbits), rng.getrandbits(pk.randbits)]
    mw, rw = to_words(ms, bits // 32), to_words(rs, pk.randbits // 32)
    out = np.zeros((len(ms), bits // 16), dtype=np.uint32)
    rc = emu.emu_encrypt_comb(shape_id(L, TPI), P(mw), bits // 32, P(rw), pk.randbits // 32, nwin, WB, P(out), bits // 16,
                              len(ms), PD(mc["n"]), U64(mc["n0inv"]), PD(nR), PD(comb), ctypes.c_uint64(len(comb)))
    assert rc == 0
    assert from_words(out) == O.encrypt_batch(pk, ms, rs)
    # make_secure = False
    rc = emu.emu_encrypt_comb(shape_id(L, TPI), P(mw), bits // 32, None, 0, nwin, WB, P(out), bits // 16,
                              len(ms), PD(mc["n"]), U64(mc["n0inv"]), PD(nR), PD(comb), ctypes.c_uint64(len(comb)))
    assert rc == 0
    assert from_words(out) == O.encrypt_batch(pk, ms, None)


def test_encrypt_finish_classic(emu):
    bits, L, TPI = 1024, 20, 2
    pk, sk = O.seeded_keypair(bits, 3, djn=False)
    N = pk.nsquare
    mc = mont_consts(N, L, TPI)
    nR = to_entry(pk.n * mc["R"] % N, L, TPI)
    rng = random.Random(10)
    ms = [0, pk.n - 1, rng.randrange(pk.n)]
    rs = [1, pk.n - 1, rng.randrange(1, pk.n)]
    obf = [O.obfuscator(pk, r) for r in rs]
    mw, ow = to_words(ms, bits // 32), to_words(obf, bits // 16)
    out = np.zeros_like(ow)
    rc = emu.emu_encrypt_finish(shape_id(L, TPI), P(mw), bits // 32, P(ow), P(out), bits // 16, len(ms), PD(mc["n"]),
                                U64(mc["n0inv"]), PD(nR), PD(mc["r2"]))
    assert rc == 0
    assert from_words(out) == O.encrypt_batch(pk, ms, rs)


@pytest.mark.parametrize("bits,L,shape", [(1024, 10, (20, 1)), (2048, 20, (20, 2)), (3072, 30, (15, 4))])
def test_decrypt_pair_engine(emu, bits, L, shape):
    """p-adic pair engine: item_dec_pair (m_p, m_q) + item_dec_crt == oracle decrypt, with the constants and the
    program exactly as the C ABI builds them for the private key (host code, no GPU)."""
    from pailliercryptolib_python_b200 import capi
    pk, sk = O.bench_keypair() if bits == 2048 else O.seeded_keypair(bits, 11)
    cpk = capi.PubKey(pk.n, bits, djn=True, hs=pk.hs)
    csk = capi.PrivKey(cpk, sk.p, sk.q)
    rng = random.Random(bits + 5)
    ms = [0, 1, pk.n - 1] + [rng.randrange(pk.n) for _ in range(3)]
    cs = [O.encrypt(pk, m, rng.getrandbits(bits // 2)) for m in ms]
    cs += [rng.randrange(1, pk.nsquare) for _ in range(2)] + [0, sk.p, sk.p * sk.q, pk.nsquare - 1]   # units and non-units
    cw = to_words(cs, bits // 16)
    hw, half = bits // 32, bits // 64
    halves = []
    for y, x in enumerate((sk.p, sk.q)):
        Lc, n0 = ctypes.c_int(), ctypes.c_uint64()
        mod = np.zeros(L, dtype=np.float64)
        cst = np.zeros(6 * 2 * L, dtype=np.float64)
        prog = np.zeros(4096, dtype=np.uint32)
        n = capi.lib().phe_privkey_pair_block(csk.h, y, ctypes.byref(Lc), ctypes.byref(n0), PD(mod), PD(cst), P(prog), len(prog))
        assert n > 0 and Lc.value == L
        # the constants are what the header says they are
        R = 1 << (52 * L)
        limbs = lambda a: sum(int(v) << (52 * i) for i, v in enumerate(a))
        assert limbs(mod) == x and n0.value == (-pow(x, -1, 1 << 52)) % (1 << 52)
        for k in range(4):
            w = (1 << (k * bits // 2)) * R * R % (x * x)
            assert (limbs(cst[(2 * k) * L:(2 * k + 1) * L]), limbs(cst[(2 * k + 1) * L:(2 * k + 2) * L])) == (w % x, w // x)
        out = np.zeros((len(cs), half), dtype=np.uint32)
        rc = emu.emu_dec_pair(L, P(cw), bits // 16, half, P(prog), P(out), half, len(cs), PD(mod), U64(n0.value), PD(cst), 32)
        assert rc == 0
        hx = sk.hp if y == 0 else sk.hq
        want = []
        for c in cs:
            u = pow(c % (x * x), x - 1, x * x)
            want.append(((u - 1) // x) * hx % x)
        assert from_words(out) == want
        halves.append(out)
        # the time-sliced form of the same program (k_dec_pair with more units than resident warps): every segment on
        # fresh lane state, only the unit's table and the parked pair carry over
        sprog, soff = capi.pair_segments(csk, y)
        assert len(soff) == 16 and soff[0] == 0
        out2 = np.zeros_like(out)
        offs = np.asarray(soff, dtype=np.int32)
        rc = emu.emu_dec_pair_segments(L, P(cw), bits // 16, half, P(sprog), len(soff), offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                       P(out2), half, len(cs), PD(mod), U64(n0.value), PD(cst), 32)
        assert rc == 0
        assert from_words(out2) == want
        # segments are balanced: no segment costs more passes (3 per multiplication, 2 per squaring) than its share,
        # except that the uncuttable preamble (conversion + window table: 4 + 1 + 31 products = 107 passes) is one piece
        cost = []
        for k in range(len(soff)):
            seg = sprog[soff[k]:(soff[k + 1] if k + 1 < len(soff) else len(sprog))]
            cost.append(sum(3 if (w & 0xff) == 7 else (2 * (int(w) >> 8) if (w & 0xff) == 8 else 0) for w in seg))
        assert sum(cost) == sum(3 if (w & 0xff) == 7 else (2 * (w >> 8) if (w & 0xff) == 8 else 0) for w in prog[:n])
        assert max(cost) <= max(107, sum(cost) / len(cost)) + 8 and min(cost) > 0, cost
    cst, n0 = _dec_consts(sk, *shape)
    mo = np.zeros((len(cs), hw), dtype=np.uint32)
    assert emu.emu_dec_crt(shape_id(*shape), P(halves[0]), P(halves[1]), half, P(mo), hw, len(cs), PD(cst), P64(n0)) == 0
    assert from_words(mo) == O.decrypt_batch(sk, cs)
    assert from_words(mo)[:len(ms)] == ms


@pytest.mark.parametrize("L,TPI,bits,block", [(20, 1, 1024, 4), (20, 2, 2048, 5), (15, 4, 3072, 4)])
def test_batched_inverse_blocks(emu, L, TPI, bits, block):
    """Montgomery's trick as run by k_inv_block: prefix products + block totals, then the unwind given the inverse of
    each total (computed here with Python ints) == element-wise pow(c, -1, N)."""
    rng = random.Random(bits + block)
    nw = bits // 32
    N = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    mc = mont_consts(N, L, TPI)
    import math
    cs = []
    while len(cs) < 3 * block - 2:
        c = rng.randrange(1, N)
        if math.gcd(c, N) == 1:
            cs.append(c)
    cs += [1, N - 1]
    cw = to_words(cs, nw)
    nb = len(cs) // block
    KP = len(mc["n"])
    P_ = np.zeros((len(cs), KP), dtype=np.float64)
    totals = np.zeros((nb, nw), dtype=np.uint32)
    rc = emu.emu_inv_block(shape_id(L, TPI), 0, P(cw), nw, len(cs), block, PD(P_), P(totals), None, None,
                           PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
    assert rc == 0
    want_tot = []
    for b in range(nb):
        t = 1
        for c in cs[b * block:(b + 1) * block]:
            t = t * c % N
        want_tot.append(t)
    assert from_words(totals) == want_tot
    tinv = to_words([pow(t, -1, N) for t in want_tot], nw)
    out = np.zeros_like(cw)
    rc = emu.emu_inv_block(shape_id(L, TPI), 1, P(cw), nw, len(cs), block, PD(P_), None, P(tinv), P(out),
                           PD(mc["n"]), U64(mc["n0inv"]), PD(mc["r2"]), PD(mc["oneM"]), PD(mc["one"]))
    assert rc == 0
    assert from_words(out) == [pow(c, -1, N) for c in cs]


# ---- n-adic pair engine (csrc/npair_items.cuh): arithmetic mod n^2 on pairs of n-sized numbers ----------------------
from emu_util import npair_consts  # noqa: E402

NPAIR_SHAPES = [(20, 1, 1024), (20, 2, 2048), (15, 4, 3072), (7, 4, 1408)]


@pytest.mark.parametrize("L,TPI,bits,win,ebits", [(20, 1, 1024, 3, 53), (20, 2, 2048, 3, 53), (20, 2, 2048, 5, 300),
                                                   (20, 2, 2048, 1, 5), (15, 4, 3072, 3, 64), (7, 4, 1408, 5, 170)])
def test_npair_mul(emu, L, TPI, bits, win, ebits):
    rng = random.Random(bits * 3 + win)
    nw = bits // 32
    ew = (ebits + 31) // 32
    for top in (bits, bits - 3):
        n = rng.getrandbits(top) | 1 | (1 << (top - 1))
        n2 = n * n
        nc = npair_consts(n, L, TPI, nw)
        base = [rng.randrange(n2) for _ in range(3)] + [0, 1, n2 - 1, n, n - 1]
        exps = [rng.getrandbits(ebits) for _ in range(3)] + [3, 0, (1 << ebits) - 1, 2, 1]
        exps[0] |= 1 << (ebits - 1)
        cw, ewa = to_words(base, 2 * nw), to_words(exps, ew)
        out = np.zeros_like(cw)
        rc = emu.emu_mul_npair(shape_id(L, TPI), win, P(cw), nw, P(ewa), ew, ew, ebits, P(out), len(base),
                               PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"]))
        assert rc == 0
        assert from_words(out) == [pow(b, e, n2) for b, e in zip(base, exps)]


def _npair_encrypt_case(emu, bits, L, TPI, WB, seed):
    rng = random.Random(seed)
    pk_o, sk_o = O.seeded_keypair(bits, seed)
    n, n2 = pk_o.n, pk_o.nsquare
    nw = bits // 32
    randbits = bits // 2
    nwin = (randbits + WB - 1) // WB
    nc = npair_consts(n, L, TPI, nw)
    KP = len(nc["cst"]) // 9
    hs_w = to_words([pk_o.hs], 2 * nw)
    comb = np.zeros(((nwin << WB) * 2 * KP,), dtype=np.float64)
    assert emu.emu_comb_npair(shape_id(L, TPI), P(hs_w), nw, nwin, WB, PD(comb), PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"])) == 0
    # spot-check table entries: T[j][d] = hs^(d 2^(WB j)) in pair-Montgomery form
    R = nc["R"]
    for j, d in [(0, 0), (0, 1), (nwin - 1, (1 << WB) - 1), (1 % nwin, 2)]:
        ent = comb[((j << WB) + d) * 2 * KP:((j << WB) + d + 1) * 2 * KP]
        x0, x1 = from_entry(ent[:KP], L, TPI), from_entry(ent[KP:], L, TPI)
        assert (x0 + x1 * n) * pow(R, -1, n2) % n2 == pow(pk_o.hs, d << (WB * j), n2)
    ms = [0, 1, n - 1, n // 3 - 1] + [rng.randrange(n) for _ in range(3)] + [rng.getrandbits(53) for _ in range(3)]
    rs = [0, 1, (1 << randbits) - 1] + [rng.getrandbits(randbits) for _ in range(len(ms) - 3)]
    mw, rw = to_words(ms, nw), to_words(rs, (randbits + 31) // 32)
    out = np.zeros((len(ms), 2 * nw), dtype=np.uint32)
    rc = emu.emu_encrypt_npair(shape_id(L, TPI), P(mw), nw, P(rw), rw.shape[1], nwin, WB, P(out), 2 * nw, len(ms),
                               PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"]), PD(comb), U64(comb.size))
    assert rc == 0
    assert from_words(out) == O.encrypt_batch(pk_o, ms, rs)
    # make_secure = False
    rc = emu.emu_encrypt_npair(shape_id(L, TPI), P(mw), nw, None, 0, nwin, WB, P(out), 2 * nw, len(ms),
                               PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"]), PD(comb), U64(comb.size))
    assert rc == 0
    assert from_words(out) == O.encrypt_batch(pk_o, ms, None)


@pytest.mark.parametrize("bits,L,TPI,WB", [(1024, 20, 1, 3), (2048, 20, 2, 2), (1408 - 64, 7, 4, 4)])
def test_npair_encrypt(emu, bits, L, TPI, WB):
    _npair_encrypt_case(emu, bits, L, TPI, WB, bits + WB)


@pytest.mark.parametrize("bits", [1024, 2048, 3072])
def test_npair_host_block_matches_python(bits):
    """The constant block the library builds for the n-adic pair engine (hostbn) == the one from Python ints, and the
    emulated HE mul on the library's block agrees with pow()."""
    from pailliercryptolib_python_b200 import capi
    pk_o, _ = O.seeded_keypair(bits, 5)
    pk = capi.PubKey(pk_o.n, bits, djn=True, hs=pk_o.hs)
    blk = capi.npair_block(pk)
    assert blk is not None
    nc = npair_consts(pk_o.n, blk["L"], blk["TPI"], bits // 32)
    assert np.array_equal(blk["cst"].reshape(-1), nc["cst"])
    assert blk["n0inv"] == nc["n0inv"] and blk["d_top"] == nc["d_top"]


@pytest.mark.parametrize("L,TPI,bits,ebits,nchunks", [(20, 1, 1024, 200, 1), (20, 2, 2048, 300, 2), (7, 4, 1408, 1408, 1)])
def test_npair_shared_exponent_program(emu, L, TPI, bits, ebits, nchunks):
    """Sliding-window program of a shared exponent on the n-adic pair engine (classic obfuscator r^n mod n^2)."""
    from pailliercryptolib_python_b200 import capi
    rng = random.Random(bits + ebits)
    nw = bits // 32
    n = rng.getrandbits(bits) | 1 | (1 << (bits - 1))
    n2 = n * n
    nc = npair_consts(n, L, TPI, nw)
    lim = n if nchunks == 1 else n2
    base = [rng.randrange(lim) for _ in range(2)] + [0, 1, lim - 1]
    bw = to_words(base, nchunks * nw)
    for e in (rng.getrandbits(ebits) | (1 << (ebits - 1)), 1 << (ebits - 1), 1, 0, 0b1000001000000):
        prog = capi.host_powm_program(e, (ebits + 31) // 32 + 1)
        pa = np.array(prog, dtype=np.uint32)
        out = np.zeros((len(base), 2 * nw), dtype=np.uint32)
        rc = emu.emu_powm_prog_npair(shape_id(L, TPI), P(bw), nw, nchunks, P(pa), len(prog) - 1, P(out), 2 * nw, len(base),
                                     PD(nc["cst"]), U64(nc["n0inv"]), U64(nc["d_top"]))
        assert rc == 0
        assert from_words(out) == [pow(b, e, n2) for b in base]
