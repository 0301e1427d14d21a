"""Executable description of the two table constructions the Snappy kernels rely on (petastorm_b200/csrc/
kernels_decode.cu: warp P of k_snappy_pages, builders/walker of k_snappy_index), checked on the CPU against a plain serial
parse of streams produced by the real compressor.  TEST INFRASTRUCTURE: a numpy model of the algorithms, not the
kernels themselves (those are compared bit for bit with pyarrow in test_gpu_decode.py)."""
import numpy as np
import pyarrow as pa
import pytest

FRAG = 65536


def _streams():
    rng = np.random.default_rng(7)
    yield 'int40', rng.integers(0, 2 ** 40, 40000, dtype=np.int64).tobytes()
    yield 'text', b' '.join(b'w%d' % (i % 977) for i in range(60000))
    yield 'runs', np.repeat(rng.integers(0, 9, 3000, dtype=np.int32), 50).tobytes()
    yield 'noise', rng.integers(0, 256, 300000, dtype=np.uint8).tobytes()
    yield 'mixed', rng.integers(0, 256, 70000, dtype=np.uint8).tobytes() + bytes(90000) + b'abc' * 30000


def _preamble(s):
    ip, shift, n = 0, 0, 0
    while True:
        b = s[ip]
        ip += 1
        n |= (b & 0x7f) << shift
        if not b & 0x80:
            return ip, n
        shift += 7


def _element(s, ip):
    """(bytes consumed, bytes produced, is_slow_path) of the element whose tag is s[ip]."""
    tag = s[ip]
    kind, t6 = tag & 3, tag >> 2
    if kind == 0:
        if t6 < 60:
            return t6 + 2, t6 + 1, False
        nb = t6 - 59
        ln = int.from_bytes(s[ip + 1:ip + 1 + nb], 'little') + 1
        return 1 + nb + ln, ln, True
    if kind == 1:
        return 2, (t6 & 7) + 4, False
    if kind == 2:
        return 3, t6 + 1, False
    return 5, t6 + 1, True


def _serial_chain(s):
    ip, ulen = _preamble(s)
    out, op = [], 0
    while ip < len(s):
        used, made, _ = _element(s, ip)
        out.append((ip, op))
        ip += used
        op += made
    assert ip == len(s) and op == ulen
    return out, ulen


def _lut():
    step = np.zeros(256, dtype=np.int64)
    made = np.zeros(256, dtype=np.int64)
    for t in range(256):
        kind, t6 = t & 3, t >> 2
        if kind == 0 and t6 < 60:
            step[t], made[t] = t6 + 2, t6 + 1
        elif kind == 1:
            step[t], made[t] = 2, (t6 & 7) + 4
        elif kind == 2:
            step[t], made[t] = 3, t6 + 1
    return step, made


@pytest.mark.parametrize('name,data', list(_streams()))
def test_successor_tables_reproduce_the_serial_chain(name, data):
    """Warp P: step[p] for every byte position of a window, quad[p] = lengths of the four elements that follow p
    (zeros stop the walk: slow-path tag, window end, stream end); walking quads visits exactly the serial positions."""
    s = pa.Codec('snappy').compress(data).to_pybytes()
    chain, _ = _serial_chain(s)
    starts = [ip for ip, _ in chain]
    step_lut, _ = _lut()
    W = 1024
    sa = np.frombuffer(s, dtype=np.uint8)
    ip, _ = _preamble(s)
    visited = []
    while ip < len(s):
        w0 = ip & ~3
        win = np.zeros(W + 64, dtype=np.int64)
        seg = sa[w0:min(w0 + W, len(s))]
        win[:len(seg)] = step_lut[seg]
        quad = np.zeros((W, 4), dtype=np.int64)
        p = np.arange(W)
        for j in range(4):                       # zeros propagate: a 0 step re-reads the same position
            quad[:, j] = win[p]
            p = p + quad[:, j]
        moved = False
        while w0 <= ip < w0 + W:
            q = quad[ip - w0]
            n = int((q != 0).sum()) if (q != 0).all() else int(np.argmax(q == 0))
            for j in range(n):
                visited.append(ip)
                ip += int(q[j])
                moved = True
            if n < 4:
                break
        if ip < len(s) and (ip < w0 + W):        # slow-path element at ip
            used, _, slow = _element(s, ip)
            assert slow
            visited.append(ip)
            ip += used
            moved = True
        assert moved
    assert visited == starts


@pytest.mark.parametrize('name,data', list(_streams()))
def test_pointer_doubling_finds_the_fragment_boundaries(name, data):
    """k_snappy_index: T[p] = {consumed, produced} of one element, four rounds of T'[p] = T[p] + T[p + consumed(T[p])]
    give 16 elements per walker step; the compressed offset at which the output position reaches every multiple of
    64 KiB equals what the serial parse finds (snappy::RawCompress never lets an element straddle a block)."""
    s = pa.Codec('snappy').compress(data).to_pybytes()
    chain, ulen = _serial_chain(s)
    expect = {op: ip for ip, op in chain if op % FRAG == 0 and op > 0}
    assert sorted(expect) == list(range(FRAG, ulen, FRAG))        # no straddling elements in compressor output
    step_lut, made_lut = _lut()
    sa = np.frombuffer(s, dtype=np.uint8)
    W = 1024
    ip, _ = _preamble(s)
    op, next_b, found = 0, FRAG, {}

    def one(ip, op, next_b):
        used, made, _ = _element(s, ip)
        if op == next_b:
            found[op] = ip
            next_b += FRAG
        assert not (op < next_b < op + made)
        return ip + used, op + made, next_b

    for j in range((len(s) + W - 1) // W):
        w0 = j * W
        cons = np.zeros(W + 64, dtype=np.int64)
        prod = np.zeros(W + 64, dtype=np.int64)
        seg = sa[w0:min(w0 + W, len(s))]
        cons[:len(seg)] = step_lut[seg]
        prod[:len(seg)] = made_lut[seg]
        for _ in range(4):
            nxt = np.arange(W) + cons[:W]
            c2, p2 = cons.copy(), prod.copy()
            c2[:W] = cons[:W] + cons[nxt]
            p2[:W] = prod[:W] + prod[nxt]
            cons, prod = c2, p2
        while w0 <= ip < min(w0 + W, len(s)):
            adv = int(cons[ip - w0])
            if adv == 0:
                ip, op, next_b = one(ip, op, next_b)          # slow-path element
                continue
            if op + int(prod[ip - w0]) >= next_b:             # a boundary lies in these elements: single steps
                end = ip + adv
                while ip < end:
                    ip, op, next_b = one(ip, op, next_b)
                continue
            ip += adv
            op += int(prod[ip - adv - w0])
    assert ip == len(s) and op == ulen
    assert found == expect
