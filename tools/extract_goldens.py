#!/usr/bin/env python3
"""Extract the reference's golden VECTORS (test data, not code) into tests/golden/ so that the
parity tests run where /root/reference does not exist (the GPU box).

  tests/golden/reference_kats.json   <- numeric vectors from
        poseidon/goldilocks_test.go:47-53        Poseidon-GL perm(0^12)
        poseidon/public_inputs_hash_test.go:53-54  HashNoPad([0,1,3736710860384812976])
        poseidon/bn254_test.go:41-88             4 Poseidon-BN254 in/out vectors
        goldilocks/quadratic_extension_test.go:27-38,70-81  QE mul / div
        goldilocks/base_test.go:108-114          MulAdd 2^63*2^63+3
        fri/fri_test.go:37-67                    8 challenger goldens on decode_block
        plonk/gates/gates_test.go:18-685,730-760 12 gates' expected constraint vectors
  tests/golden/testdata/{step,decode_block}/*.json   <- the two real Plonky2 proof fixtures (byte copies)

Run in the build container only; outputs are committed.
"""
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests/golden")


def qe_list(body):
    return [[a, b] for a, b in re.findall(r'NewVariable\("(\d+)"\), gl\.NewVariable\("(\d+)"\)', body)]


def main():
    os.makedirs(OUT, exist_ok=True)
    k = {}
    # ---- gates -------------------------------------------------------------------------
    src = open(os.path.join(REF, "plonk/gates/gates_test.go")).read()
    gates = {}
    for m in re.finditer(r"var (\w+) = \[\]gl\.QuadraticExtensionVariable\{(.*?)\n\}", src, re.S):
        gates[m.group(1)] = qe_list(m.group(2))
    assert len(gates["localConstants"]) == 5 and len(gates["localWires"]) == 136
    bw = re.search(r"NewCosetInterpolationGate\(\s*(\d+),\s*(\d+),\s*\[\]goldilocks\.Element\{(.*?)\}", src, re.S)
    weights = re.findall(r"NewElement\((\d+)\)", bw.group(3))
    tests = []
    for m in re.finditer(r"\{(?:gates\.New(\w+)\(([^)]*)\)|&gates\.(\w+)\{\}), (\w+)\},", src):
        name = m.group(1) or m.group(3)
        args = [int(x) for x in re.findall(r"\d+", m.group(2) or "")]
        tests.append({"gate": name, "args": args, "expected": m.group(4)})
    tests.append({"gate": "CosetInterpolationGate", "args": [int(bw.group(1)), int(bw.group(2))],
                  "weights": weights, "expected": "cosetInterpolationGateExpectedConstraints"})
    assert len(tests) == 11, tests
    k["gates"] = {"vectors": gates, "tests": tests, "public_inputs_hash": ["0", "0", "0", "0"],
                  "common_data": "decode_block"}
    # ---- poseidon GL ---------------------------------------------------------------------
    src = open(os.path.join(REF, "poseidon/goldilocks_test.go")).read()
    m = re.search(r"outStr := \[\]string\{(.*?)\}", src, re.S)
    k["poseidon_gl_perm_zero"] = re.findall(r'"(\d+)"', m.group(1))
    src = open(os.path.join(REF, "poseidon/public_inputs_hash_test.go")).read()
    k["public_inputs_hash"] = {
        "in": re.findall(r'"(\d+)"', re.search(r"inStr := \[\]string\{(.*?)\}", src).group(1)),
        "out": re.findall(r'"(\d+)"', re.search(r"outStr := \[\]string\{(.*?)\}", src).group(1))}
    # ---- poseidon BN254 ----------------------------------------------------------------
    src = open(os.path.join(REF, "poseidon/bn254_test.go")).read()
    body = re.search(r"testCases := \[\]\[2\]\[\]string\{(.*?)\n\t\}\n", src, re.S).group(1)
    nums = re.findall(r'"(\d+)"', body)
    assert len(nums) == 32
    k["poseidon_bn254"] = [{"in": nums[8 * i:8 * i + 4], "out": nums[8 * i + 4:8 * i + 8]} for i in range(4)]
    # ---- QE ------------------------------------------------------------------------------
    src = open(os.path.join(REF, "goldilocks/quadratic_extension_test.go")).read()
    nums = re.findall(r'NewVariable\("(\d+)"\)', src)
    assert len(nums) == 12
    k["qe_mul"] = {"a": nums[0:2], "b": nums[2:4], "out": nums[4:6]}
    k["qe_div"] = {"a": nums[6:8], "b": nums[8:10], "out": nums[10:12]}
    # ---- MulAdd ----------------------------------------------------------------------------
    src = open(os.path.join(REF, "goldilocks/base_test.go")).read()
    m = re.search(r'SetString\("(\d+)", 10\)', src)
    k["muladd"] = {"a": str(1 << 63), "b": str(1 << 63), "c": "3", "out": m.group(1)}
    assert "1 << 63" in src or "9223372036854775808" in src
    # ---- challenger goldens on decode_block -----------------------------------------------
    src = open(os.path.join(REF, "fri/fri_test.go")).read()
    nv = re.findall(r'gl\.NewVariable\("(\d+)"\)\)', src)
    lim = re.findall(r"\.Limb, (\d+)\)", src)
    x = re.search(r"x = (\d+)", src).group(1)
    k["challenger_decode_block"] = {"plonk_beta0": nv[0], "plonk_gamma0": nv[1], "plonk_alpha0": nv[2],
                                    "plonk_zeta0": nv[3], "fri_alpha0": lim[0], "fri_beta00": lim[1],
                                    "fri_pow_response": lim[2], "fri_query_index0": x}
    json.dump(k, open(os.path.join(OUT, "reference_kats.json"), "w"), indent=0)
    for d in ("step", "decode_block"):
        dst = os.path.join(OUT, "testdata", d)
        os.makedirs(dst, exist_ok=True)
        for f in ("common_circuit_data.json", "proof_with_public_inputs.json", "verifier_only_circuit_data.json"):
            shutil.copyfile(os.path.join(REF, "testdata", d, f), os.path.join(dst, f))
    print("ok:", [t["gate"] for t in tests], k["challenger_decode_block"])


if __name__ == "__main__":
    main()
