"""Gate vectors the reference does not have (plonk/gates/gates_test.go covers 11 of the 14 gates; ExponentiationGate,
ConstantGate and NoopGate are only exercised through "the fixture's proof verifies").

The C++ gadgets (csrc/host/gadgets_fri_plonk.cc) and the Python oracle (oracle/gates.py) both restate the reference's Go
statement by statement, so a shared misreading would pass in both. This file evaluates the gates a SECOND way, straight from
Plonky2's definition of each gate's constraint polynomial, with its own three-line GF(p^2) arithmetic:

  ConstantGate{n}:        c_i - w_i                                                  (i < n)
  NoopGate:               no constraints
  ExponentiationGate{n}:  wires = base, n power bits (little endian), output, n intermediate values;
                          prev_0 = 1, prev_i = t_{i-1}^2, bit_i = bits[n-1-i] (most significant first),
                          constraint_i = prev_i * (bit_i * base + 1 - bit_i) - t_i,  last = output - t_{n-1}
                          -> on a witness built by square-and-multiply every constraint is 0 and output = base^power.

and checks (1) the oracle against it on random wires, (2) that a square-and-multiply witness zeroes every constraint and a
flipped bit does not, (3) the C++ gadget, run as a one-gate circuit through the test-only host interpreter, against it, and
(4) the harness itself on a gate that DOES have a reference vector (ArithmeticGate, plonk/gates/gates_test.go:712-768).
The same one-gate circuits run on the GPU in tests/test_gpu_wrap.py::test_gate_circuits_on_gpu."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import gpw
from oracle import gates as og
from oracle import goldilocks as ogl
from oracle.engine import Api
from oracle.types import read_common_circuit_data

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTLIB = os.path.join(ROOT, "tests", "hostlib")
P = (1 << 64) - (1 << 32) + 1
EXP_ID = "ExponentiationGate { num_power_bits: %d, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>"


# ---- independent GF(p^2) = GF(p)[x] / (x^2 - 7) ---------------------------------------------------------------------------
def qadd(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def qsub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def qmul(a, b):
    return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


ONE, ZERO = (1, 0), (0, 0)


def exponentiation_constraints(n, wires):
    base, bits, output, inter = wires[0], wires[1:1 + n], wires[1 + n], wires[2 + n:2 + 2 * n]
    out = []
    for i in range(n):
        prev = ONE if i == 0 else qmul(inter[i - 1], inter[i - 1])
        bit = bits[n - 1 - i]
        out.append(qsub(qmul(prev, qadd(qmul(bit, base), qsub(ONE, bit))), inter[i]))
    out.append(qsub(output, inter[n - 1]))
    return out


def exponentiation_witness(n, base, power):
    bits = [((power >> i) & 1, 0) for i in range(n)]
    inter, acc = [], ONE
    for i in range(n):
        acc = qmul(acc, acc)
        if (power >> (n - 1 - i)) & 1:
            acc = qmul(acc, base)
        inter.append(acc)
    return [base] + bits + [inter[-1]] + inter


def constant_constraints(n, consts, wires):
    return [qsub(consts[i], wires[i]) for i in range(n)]


def _rand_qe(rng):
    return (rng.randrange(P), rng.randrange(P))


def _oracle_eval(gate, consts, wires, pih=(0, 0, 0, 0)):
    api = Api(trace=False)
    return [tuple(int(x) for x in c) for c in gate.EvalUnfiltered(api, ogl.Chip(api), consts, wires, list(pih))]


# ---- (1) + (2): the oracle against the second form ------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 5, 67])
def test_exponentiation_gate_oracle_vs_definition(n):
    rng = random.Random(n)
    gate = og.GateInstanceFromId(EXP_ID % n)
    for _ in range(3):  # arbitrary (unsatisfying) wires: both forms must give the same non-zero values
        wires = [_rand_qe(rng) for _ in range(2 + 2 * n)]
        assert _oracle_eval(gate, [], wires) == exponentiation_constraints(n, wires)
    base, power = _rand_qe(rng), rng.randrange(1 << n)
    wires = exponentiation_witness(n, base, power)
    # square-and-multiply really computes base^power in GF(p^2)
    acc = ONE
    for _ in range(power if n <= 5 else 0):
        acc = qmul(acc, base)
    if n <= 5:
        assert wires[1 + n] == acc
    assert _oracle_eval(gate, [], wires) == [ZERO] * (n + 1)
    wires[1] = qsub(ONE, wires[1])  # flip the least significant power bit: it is consumed by the LAST step
    got = _oracle_eval(gate, [], wires)
    assert got == exponentiation_constraints(n, wires)
    assert got[n - 1] != ZERO and all(c == ZERO for i, c in enumerate(got) if i != n - 1)


def test_constant_and_noop_gate_oracle_vs_definition():
    rng = random.Random(5)
    for n in (1, 2, 4):
        gate = og.GateInstanceFromId("ConstantGate { num_consts: %d }" % n)
        consts = [_rand_qe(rng) for _ in range(n + 1)]
        wires = [_rand_qe(rng) for _ in range(n + 3)]
        assert _oracle_eval(gate, consts, wires) == constant_constraints(n, consts, wires)
        assert _oracle_eval(gate, consts, consts) == [ZERO] * n
    assert _oracle_eval(og.GateInstanceFromId("NoopGate"), [_rand_qe(rng)], [_rand_qe(rng)] * 3) == []


# ---- (3) + (4): the C++ gadget as a one-gate circuit --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def lib():
    subprocess.check_call(["make", "-C", HOSTLIB], stdout=subprocess.DEVNULL)
    lib = C.CDLL(os.path.join(HOSTLIB, "libgpw_circuit_test.so"))
    lib.ct_compile_gate.restype = C.c_void_p
    lib.ct_compile_gate.argtypes = [C.c_char_p]
    lib.ct_free.argtypes = [C.c_void_p]
    lib.ct_last_error.restype = C.c_char_p
    lib.ct_solve_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ct_check.restype = C.c_uint64
    lib.ct_check.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def gate_circuit_inputs(expected, consts, wires, pih=(0, 0, 0, 0)):
    """(public, secret) integer lists in DefineGateCircuit order"""
    pub = [x for c in expected for x in c]
    sec = [x for c in consts for x in c] + [x for w in wires for x in w] + list(pih)
    return pub, sec


def _gadget_unsatisfied(lib, spec, expected, consts, wires, pih=(0, 0, 0, 0)):
    h = lib.ct_compile_gate(spec.encode())
    assert h, lib.ct_last_error()
    pub, sec = gate_circuit_inputs(expected, consts, wires, pih)
    pl, sl, xl = gpw.ints_to_limbs(pub), gpw.ints_to_limbs(sec), gpw.ints_to_limbs([0x1234567890abcdef1234567890abcdef])
    rc = lib.ct_solve_inputs(h, pl.ctypes.data, len(pub), sl.ctypes.data, len(sec), xl.ctypes.data)
    assert rc == 0, lib.ct_last_error()
    fb = C.c_int64()
    bad = lib.ct_check(h, C.byref(fb))
    lib.ct_free(h)
    return bad


def gate_cases(rng):
    """(spec, expected constraint values by the second form, consts, wires) for the three vector-less gates"""
    cases = []
    for n in (3, 67):
        wires = [_rand_qe(rng) for _ in range(2 + 2 * n)]
        cases.append(("0:%d:%d:%s" % (len(wires), n + 1, EXP_ID % n), exponentiation_constraints(n, wires), [], wires))
        wires = exponentiation_witness(n, _rand_qe(rng), rng.randrange(1 << n))
        cases.append(("0:%d:%d:%s" % (len(wires), n + 1, EXP_ID % n), [ZERO] * (n + 1), [], wires))
    consts, wires = [_rand_qe(rng) for _ in range(2)], [_rand_qe(rng) for _ in range(4)]
    cases.append(("2:4:2:ConstantGate { num_consts: 2 }", constant_constraints(2, consts, wires), consts, wires))
    cases.append(("1:2:0:NoopGate", [], [_rand_qe(rng)], [_rand_qe(rng)] * 2))
    return cases


def test_cpp_gate_gadgets_vs_definition(lib):
    rng = random.Random(21)
    for spec, expected, consts, wires in gate_cases(rng):
        assert _gadget_unsatisfied(lib, spec, expected, consts, wires) == 0, spec
        if expected:
            wrong = [qadd(expected[0], ONE)] + expected[1:]
            assert _gadget_unsatisfied(lib, spec, wrong, consts, wires) > 0, spec


def test_harness_on_a_gate_with_a_reference_vector(lib, kats, testdata_dir):
    # ArithmeticGate on the reference's own vector (plonk/gates/gates_test.go): the one-gate circuit accepts the reference's
    # expected constraints and rejects a perturbed one - so "satisfied" above means what it should
    g = kats["gates"]
    common = read_common_circuit_data(os.path.join(testdata_dir, g["common_data"], "common_circuit_data.json"))
    qe = lambda v: [(int(a), int(b)) for a, b in v]
    consts = qe(g["vectors"]["localConstants"])[len(common.SelectorGroups):]
    wires = qe(g["vectors"]["localWires"])
    pih = [int(x) for x in g["public_inputs_hash"]]
    t = next(t for t in g["tests"] if t["gate"] == "ArithmeticGate")
    expected = qe(g["vectors"][t["expected"]])
    spec = "%d:%d:%d:ArithmeticGate { num_ops: %d }" % (len(consts), len(wires), len(expected), t["args"][0])
    assert _gadget_unsatisfied(lib, spec, expected, consts, wires, pih) == 0
    wrong = expected[:-1] + [qadd(expected[-1], ONE)]
    assert _gadget_unsatisfied(lib, spec, wrong, consts, wires, pih) > 0
