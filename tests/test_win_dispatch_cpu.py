"""The generated inline-PTX dispatch of the register-window gather
(csrc/win_dispatch.cuh, tools/gen_win_dispatch.py) interpreted on the CPU.

The header is a few thousand lines of machine-written PTX (jump tables, one run of 8
fma per delay offset, operand numbers into a 28- or 92-operand asm statement).  A tiny
interpreter for exactly the instructions it uses executes every specialisation on
random inputs and compares with the definition
    acc[s][k] += w[s] * win[k + W - rel[s]]      (rel[s] <= W; 255 = no pair).
This pins operand numbering, table contents and fall-through structure; it does not
(and cannot) check how ptxas assembles the text."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "sparrowpy_b200", "csrc", "win_dispatch.cuh")


def parse_functions():
    text = open(HEADER).read()
    funcs = {}
    for m in re.finditer(r"void (accumulate_\w+)<([\d, ]+)>\(.*?asm volatile\((.*?)\);\n}",
                         text, flags=re.S):
        name, targs, body = m.group(1), m.group(2).replace(" ", ""), m.group(3)
        lines = re.findall(r'"(.*?)\\n"', body)
        n_out = body.count('"+d"(')
        n_in = body.count('"d"(') + body.count('"r"(')
        funcs[(name, targs)] = (lines, n_out, n_in)
    return funcs


def interpret(lines, ops):
    """Execute the PTX subset on the operand list ``ops`` (floats and ints, modified in
    place); returns the number of fma executed."""
    tables, labels, prog = {}, {}, []
    for ln in lines:
        ln = ln.strip().rstrip(";")
        if ln in ("{", "}") or ln.startswith(".reg"):
            continue
        m = re.match(r"(\w+): \.branchtargets (.*)", ln)
        if m:
            tables[m.group(1)] = [x.strip() for x in m.group(2).split(",")]
            continue
        m = re.match(r"(\w+):$", ln)
        if m:
            labels[m.group(1)] = len(prog)
            continue
        prog.append(ln)
    val = lambda tok: ops[int(tok[1:])] if tok.startswith("%") else int(tok)   # noqa: E731
    pc, idx, n_fma, steps = 0, 0, 0, 0
    while pc < len(prog):
        steps += 1
        assert steps < 10000, "runaway"
        op, _, rest = prog[pc].partition(" ")
        args = [a.strip() for a in rest.split(",")]
        pc += 1
        if op == "fma.rn.f64":
            d, a, b, c = args
            assert d == c
            ops[int(d[1:])] = float(np.float64(val(a)) * np.float64(val(b)) + ops[int(c[1:])])
            n_fma += 1
        elif op == "and.b32":
            idx = val(args[1]) & val(args[2])
        elif op == "bfe.u32":
            idx = (val(args[1]) >> val(args[2])) & ((1 << val(args[3])) - 1)
        elif op == "brx.idx.uni":
            assert len(tables[args[1]]) == 16 and 0 <= idx < 16
            pc = labels[tables[args[1]][idx]]
        elif op == "bra.uni":
            pc = labels[args[0]]
        else:
            raise AssertionError("unexpected instruction " + prog[pc - 1])
    return n_fma


def test_header_is_up_to_date(tmp_path):
    """the committed header is what the generator produces"""
    fresh = str(tmp_path / "win_dispatch.cuh")
    subprocess.check_call([sys.executable, os.path.join(REPO, "tools", "gen_win_dispatch.py"),
                           fresh], stdout=subprocess.DEVNULL)
    assert open(fresh).read() == open(HEADER).read()


@pytest.mark.parametrize("width", [4, 10])
def test_per_receiver_dispatch(width):
    lines, n_out, n_in = parse_functions()[("accumulate_brx", f"{width},8")]
    assert (n_out, n_in) == (8, 8 + width + 2)
    rng = np.random.default_rng(width)
    for rel in list(range(width + 1)) + [255]:
        acc, win, w = rng.normal(size=8), rng.normal(size=8 + width), float(rng.normal())
        ops = list(acc) + list(win) + [w, rel]
        n_fma = interpret(lines, ops)
        want = acc + w * win[width - rel: width - rel + 8] if rel <= width else acc
        assert n_fma == (8 if rel <= width else 0)
        assert np.array_equal(np.array(ops[:8]), want)


@pytest.mark.parametrize("width", [4, 10])
def test_chained_record_dispatch(width):
    lines, n_out, n_in = parse_functions()[("accumulate_chain", str(width))]
    assert (n_out, n_in) == (64, 8 + width + 8 + 2)
    rng = np.random.default_rng(100 + width)
    for trial in range(40):
        acc = rng.normal(size=(8, 8))
        win, w = rng.normal(size=8 + width), rng.normal(size=8)
        rel = rng.integers(0, width + 1, size=8)
        rel[rng.random(8) < 0.25] = 255                       # empty slots
        if trial == 0:
            rel[:] = 255
        packed = int.from_bytes(bytes(int(r) for r in rel), "little")
        ops = list(acc.reshape(-1)) + list(win) + list(w) + [packed & 0xffffffff, packed >> 32]
        n_fma = interpret(lines, ops)
        want = acc.copy()
        for s in range(8):
            if rel[s] <= width:
                want[s] += w[s] * win[width - rel[s]: width - rel[s] + 8]
        assert n_fma == 8 * int((rel <= width).sum())
        assert np.array_equal(np.array(ops[:64]).reshape(8, 8), want)
