// TEST INFRASTRUCTURE ONLY (built into tests/hostlib/libgpw_circuit_test.so; never linked into libgpw.so and not
// reachable from any product entry point): a sequential host interpreter of the solver tape, used to validate
// the C++ gadget library and the frontend on machines without a GPU, and as the reference the CUDA tape
// executor is compared against wire-for-wire. It is NOT a fallback: the product solves witnesses on the GPU only.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../gnark-plonky2-verifier_b200/csrc/gl.cuh"
#include "../../gnark-plonky2-verifier_b200/csrc/host/frontend.h"
#include "../../gnark-plonky2-verifier_b200/csrc/host/gadgets.h"
#include "../../gnark-plonky2-verifier_b200/csrc/host/scs.h"
#include "../../gnark-plonky2-verifier_b200/csrc/poseidon_bn254_macro.cuh"
#include "../../gnark-plonky2-verifier_b200/csrc/poseidon_gl_macro.cuh"
#include "../../gnark-plonky2-verifier_b200/csrc/poseidon_constants.inc"

using namespace gpw;
using namespace gpw::fe;

struct Circuit {
  API api;
  gadgets::CommonCircuitData cd;
  std::vector<Fr> w;
  std::string err;
  size_t n_baked = 0;
};

static thread_local std::string g_err;

static Fr eval_le(const Circuit& c, uint32_t le) {
  const auto& off = c.api.LeOffsets();
  const auto& wi = c.api.LeWires();
  const auto& ci = c.api.LeCoeffIds();
  const auto& co = c.api.Coeffs();
  Fr acc = Fr::zero();
  for (uint32_t k = off[le]; k < off[le + 1]; k++) {
    const Fr& v = c.w[wi[k]];
    if (ci[k] == API::COEFF_ONE) acc = add(acc, v);
    else if (ci[k] == API::COEFF_NEG_ONE) acc = sub(acc, v);
    else acc = add(acc, mul(co[ci[k]], v));
  }
  return acc;
}

static void canon(const Fr& a, uint64_t l[4]) { fr_to_limbs(a, l); }

extern "C" {

const char* ct_last_error() { return g_err.c_str(); }

void* ct_compile(const char* common_json) {
  try {
    Circuit* c = new Circuit();
    c->cd = gadgets::ReadCommonCircuitData(common_json);
    gadgets::DefineVerifierCircuit(&c->api, c->cd);
    return c;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// PLONK lowering (host/scs.cc) of the solved circuit: out6 = {gates, variables, logN, unsatisfied gates, first bad row + 1,
// permutation errors}. The witness is extended on the host, every gate equation evaluated, and sigma checked to be a
// permutation of the 3 N slots that never leaves a variable's slots.
void ct_scs_check(void* h, uint64_t* out6) {
  Circuit* c = (Circuit*)h;
  scs::System sys = scs::Build(c->api);
  std::vector<Fr> v(c->w.begin(), c->w.end());
  scs::ExtendWitness(sys, &v);
  int64_t fb = -1;
  const uint64_t bad = scs::CheckGates(sys, v, &fb);
  std::vector<uint32_t> sigma;
  scs::BuildPermutation(sys, &sigma);
  const size_t N = (size_t)1 << sys.logN;
  auto var_of = [&](size_t slot) -> uint32_t {
    const size_t col = slot / N, row = slot % N;
    if (row >= sys.n_gates) return 0;
    return col == 0 ? sys.a[row] : col == 1 ? sys.b[row] : sys.c[row];
  };
  uint64_t perm_bad = 0;
  std::vector<uint8_t> hit(3 * N, 0);
  for (size_t t = 0; t < 3 * N; t++) {
    if (sigma[t] >= 3 * N || hit[sigma[t]]) {
      perm_bad++;
      continue;
    }
    hit[sigma[t]] = 1;
    if (var_of(sigma[t]) != var_of(t)) perm_bad++;
  }
  out6[0] = sys.n_gates;
  out6[1] = sys.n_vars;
  out6[2] = (uint64_t)sys.logN;
  out6[3] = bad;
  out6[4] = (uint64_t)(fb + 1);
  out6[5] = perm_bad;
}

// compile cache round trip of the host-side compiled circuit (fe::API::Serialize / Deserialize)
int ct_save(void* h, const char* path) {
  std::ofstream os(path, std::ios::binary | std::ios::trunc);
  if (!os) return -1;
  ((Circuit*)h)->api.Serialize(os);
  return os ? 0 : -1;
}
void* ct_load(const char* path) {
  try {
    std::ifstream is(path, std::ios::binary);
    if (!is) throw std::runtime_error("cannot open file");
    Circuit* c = new Circuit();
    c->api.Deserialize(is);
    return c;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// one gate as a circuit (gadgets::DefineGateCircuit): spec = "n_consts:n_wires:n_constraints:gate id"
void* ct_compile_gate(const char* spec) {
  try {
    Circuit* c = new Circuit();
    gadgets::DefineGateCircuit(&c->api, spec);
    return c;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// bound / baked circuit forms (verifier/util.go:10-24): mode 1 = verifier-only data constant, 2 = proof constant too
void* ct_compile_baked(const char* common_json, const char* proof_json, const char* vo_json, int mode) {
  try {
    Circuit* c = new Circuit();
    c->cd = gadgets::ReadCommonCircuitData(common_json);
    std::vector<std::array<uint64_t, 4>> baked =
        mode == 2 ? gadgets::ParseProofInputs(c->cd, proof_json, vo_json).sec : gadgets::ParseVerifierOnly(c->cd, vo_json);
    c->n_baked = baked.size();
    gadgets::DefineVerifierCircuit(&c->api, c->cd, &baked);
    return c;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

// Small self-contained circuits for unit tests of the gadget port. kind: 0 = Poseidon-GL permutation of 12 secret
// inputs, outputs asserted equal to 12 public inputs; 1 = Poseidon-BN254 (4 secret in, 4 public out);
// 2 = QE mul + div (secret a, b; public out a*b and a/b); 3 = RangeCheck of one secret input.
void* ct_compile_small(int kind) {
  try {
    Circuit* c = new Circuit();
    API* api = &c->api;
    gadgets::GlChip gl(api);
    if (kind == 0) {
      std::vector<Variable> out, in;
      for (int i = 0; i < 12; i++) out.push_back(api->PublicInput());
      for (int i = 0; i < 12; i++) in.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::PoseidonGlChip p(api);
      gadgets::GlState st;
      for (int i = 0; i < 12; i++) st[i] = in[i];
      st = p.Poseidon(st);
      for (int i = 0; i < 12; i++) api->AssertIsEqual(st[i], out[i]);
    } else if (kind == 1) {
      std::vector<Variable> out, in;
      for (int i = 0; i < 4; i++) out.push_back(api->PublicInput());
      for (int i = 0; i < 4; i++) in.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::PoseidonBn254Chip p(api);
      auto st = p.Poseidon({in[0], in[1], in[2], in[3]});
      for (int i = 0; i < 4; i++) api->AssertIsEqual(st[i], out[i]);
    } else if (kind == 2) {
      std::vector<Variable> out, in;
      for (int i = 0; i < 4; i++) out.push_back(api->PublicInput());
      for (int i = 0; i < 4; i++) in.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::QE a = {in[0], in[1]}, b = {in[2], in[3]};
      gadgets::QE m = gl.MulExtension(a, b);
      auto d = gl.DivExtension(a, b);
      api->AssertIsEqual(m[0], out[0]);
      api->AssertIsEqual(m[1], out[1]);
      api->AssertIsEqual(d.first[0], out[2]);
      api->AssertIsEqual(d.first[1], out[3]);
    } else if (kind == 3) {
      Variable x = api->SecretInput();
      api->EndInputs();
      gl.RangeCheck(x);
    } else {
      throw std::runtime_error("unknown small circuit kind");
    }
    api->Finalize();
    return c;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void ct_schedule_alap(void* h) { ((Circuit*)h)->api.ScheduleALAP(); }
void ct_schedule_spine_tail(void* h) { ((Circuit*)h)->api.ScheduleSpineAndTail(); }

// Structural check of the (scheduled) tape the GPU executor runs: every input wire of an instruction is a circuit input
// or is produced by an instruction of a strictly LOWER level; every wire has at most one producer; macro instructions
// carry as many output wires as they claim. out3 = {level violations, doubly produced wires, macro instructions}.
void ct_schedule_check(void* h, uint64_t* out3) {
  Circuit* c = (Circuit*)h;
  const API& api = c->api;
  const auto& tape = api.Tape();
  const auto& mo = api.MacroOuts();
  const auto& off = api.LeOffsets();
  const auto& wi = api.LeWires();
  std::vector<uint32_t> producer_level(api.NumWires(), 0xffffffffu);
  uint64_t bad_level = 0, dup = 0, macros = 0;
  for (const auto& in : tape) {
    if (in.outs_off != NO_LE) macros++;
    for (uint32_t k = 0; k < in.nout; k++) {
      const uint32_t w = in.out_wire(k, mo);
      if (producer_level[w] != 0xffffffffu) dup++;
      producer_level[w] = in.level;
    }
  }
  for (const auto& in : tape) {
    const uint32_t les[4] = {in.le[0], in.le[1], in.le[2], in.le3};
    for (uint32_t le : les) {
      if (le == NO_LE) continue;
      for (uint32_t t = off[le]; t < off[le + 1]; t++) {
        const uint32_t pl = producer_level[wi[t]];
        if (pl != 0xffffffffu && pl >= in.level) {
          if (bad_level < 8 && getenv("CT_DEBUG"))
            fprintf(stderr, "violation: op %d level %u out %u reads wire %u produced at level %u\n", in.op, in.level, in.out, wi[t], pl);
          bad_level++;
        }
      }
    }
  }
  out3[0] = bad_level;
  out3[1] = dup;
  out3[2] = macros;
}

void ct_free(void* h) { delete (Circuit*)h; }

// stats: [wires, public, secret, constraints, tape, levels, commit_level, limb_wires, muladd, reduce, glinv, split,
//         invzero, bits, div, decomp, mul, coeffs, le_terms, limb_wire_start, count_wire_start, commit_wire]
void ct_stats(void* h, uint64_t* out) {
  Circuit* c = (Circuit*)h;
  const auto& k = c->api.Counts();
  uint64_t v[] = {c->api.NumWires(), c->api.NumPublic(), c->api.NumSecret(), c->api.NumConstraints(), c->api.Tape().size(),
                  c->api.NumLevels(), c->api.CommitLevel(), c->api.NumLimbWires(), k.muladd, k.reduce, k.glinv, k.split,
                  k.invzero, k.bits, k.div, k.decomp, k.mul, c->api.Coeffs().size(), c->api.LeWires().size(),
                  c->api.LimbWireStart(), c->api.CountWireStart(), c->api.CommitWire()};
  memcpy(out, v, sizeof(v));
}

// number of tape instructions per level (out has NumLevels entries)
void ct_level_histogram(void* h, uint32_t* out) {
  Circuit* c = (Circuit*)h;
  for (uint32_t i = 0; i < c->api.NumLevels(); i++) out[i] = 0;
  for (const auto& in : c->api.Tape()) out[in.level]++;
}

// inputs: canonical limbs. Returns 0 ok, <0 on hint precondition failure. x_commit: challenge used for OP_COMMIT.
int ct_solve_inputs(void* h, const uint64_t* pub, size_t npub, const uint64_t* sec, size_t nsec, const uint64_t* x_commit) {
  Circuit* c = (Circuit*)h;
  API& api = c->api;
  if (npub != api.NumPublic() || nsec != api.NumSecret()) {
    g_err = "input count mismatch: circuit wants " + std::to_string(api.NumPublic()) + " public / " +
            std::to_string(api.NumSecret()) + " secret";
    return -1;
  }
  c->w.assign(api.NumWires(), Fr::zero());
  c->w[0] = Fr::one();
  for (size_t i = 0; i < npub; i++) c->w[1 + i] = fr_from_limbs(pub + 4 * i);
  for (size_t i = 0; i < nsec; i++) c->w[1 + npub + i] = fr_from_limbs(sec + 4 * i);
  std::vector<uint32_t> hist;
  // the post-commitment DIVs are independent of each other: invert their denominators with Montgomery's trick
  std::vector<const Instr*> divs;
  for (const auto& in : api.Tape()) {
    if (in.op == OP_DIV && in.level > api.CommitLevel() && api.CommitLevel() != 0) {
      divs.push_back(&in);
      continue;
    }
    switch (in.op) {
      case OP_MUL: {
        Fr r = mul(eval_le(*c, in.le[0]), eval_le(*c, in.le[1]));
        if (in.le[2] != NO_LE) r = add(r, eval_le(*c, in.le[2]));
        c->w[in.out] = r;
        break;
      }
      case OP_HINT_MULADD: {
        uint64_t a[4], b[4], d[4];
        canon(eval_le(*c, in.le[0]), a);
        canon(eval_le(*c, in.le[1]), b);
        canon(eval_le(*c, in.le[2]), d);
        if (a[1] | a[2] | a[3] | b[1] | b[2] | b[3] | d[1] | d[2] | d[3] || a[0] >= gl::P || b[0] >= gl::P || d[0] >= gl::P) {
          g_err = "MulAddHint: operand is not in the field";  // base.go:228-232
          return -5;
        }
        uint64_t q, r;
        gl::mul_add_hint(a[0], b[0], d[0], q, r);
        c->w[in.out] = fr_from_u64(q);
        c->w[in.out + 1] = fr_from_u64(r);
        break;
      }
      case OP_HINT_REDUCE: {
        uint64_t x[4], q[4], r;
        canon(eval_le(*c, in.le[0]), x);
        gl::reduce_hint(x, q, r);
        c->w[in.out] = fr_from_limbs(q);
        c->w[in.out + 1] = fr_from_u64(r);
        break;
      }
      case OP_HINT_GLINV: {
        uint64_t x[4];
        canon(eval_le(*c, in.le[0]), x);
        if (x[1] | x[2] | x[3] || x[0] >= gl::P) {
          g_err = "InverseHint: input is not in the field";
          return -5;
        }
        c->w[in.out] = fr_from_u64(gl::inverse(x[0]));
        break;
      }
      case OP_HINT_SPLIT: {
        uint64_t x[4];
        canon(eval_le(*c, in.le[0]), x);
        if (x[1] | x[2] | x[3] || x[0] >= gl::P) {
          g_err = "SplitLimbsHint: input is not in the field";
          return -5;
        }
        c->w[in.out] = fr_from_u64(x[0] >> 32);
        c->w[in.out + 1] = fr_from_u64(x[0] & 0xffffffffull);
        break;
      }
      case OP_INVZERO: c->w[in.out] = inv(eval_le(*c, in.le[0])); break;
      case OP_BITS: {
        uint64_t x[4];
        canon(eval_le(*c, in.le[0]), x);
        for (uint32_t i = 0; i < in.nout; i++) c->w[in.out + i] = fr_from_u64((x[i >> 6] >> (i & 63)) & 1);
        break;
      }
      case OP_DIV: c->w[in.out] = mul(eval_le(*c, in.le[0]), inv(eval_le(*c, in.le[1]))); break;
      case OP_DECOMP: {
        uint64_t x[4];
        canon(eval_le(*c, in.le[0]), x);
        for (uint32_t i = 0; i < in.nout; i++) c->w[in.out + i] = fr_from_u64((x[(16 * i) >> 6] >> ((16 * i) & 63)) & 0xffff);
        break;
      }
      case OP_COUNT: {
        hist.assign(65536, 0);
        for (uint32_t i = 0; i < api.NumLimbWires(); i++) {
          uint64_t x[4];
          canon(c->w[api.LimbWireStart() + i], x);
          hist[x[0] & 0xffff]++;
        }
        for (uint32_t i = 0; i < 65536; i++) c->w[in.out + i] = fr_from_u64(hist[i]);
        break;
      }
      case OP_COMMIT: c->w[in.out] = fr_from_limbs(x_commit); break;
      case OP_POSEIDON_BN254: {
        const uint32_t les[4] = {in.le[0], in.le[1], in.le[2], in.le3};
        const auto& off = api.LeOffsets();
        const auto& wi = api.LeWires();
        Fr st[4];
        bool isc[4];
        for (int k = 0; k < 4; k++) {
          st[k] = eval_le(*c, les[k]);
          uint32_t n = off[les[k] + 1] - off[les[k]];
          isc[k] = n == 0 || (n == 1 && wi[off[les[k]]] == 0);
        }
        Bn254PoseidonTables T{reinterpret_cast<const Fr*>(GPW_BN_C_MONT), reinterpret_cast<const Fr*>(GPW_BN_S_MONT),
                              reinterpret_cast<const Fr*>(GPW_BN_M_MONT), reinterpret_cast<const Fr*>(GPW_BN_P_MONT)};
        uint32_t idx = 0;
        poseidon_bn254_trace(st, isc, T, [&](const Fr& v) { c->w[in.out + idx++] = v; });
        if (idx != in.nout) { g_err = "poseidon macro emitted a different number of wires than the builder created"; return -1; }
        break;
      }
      case OP_POSEIDON_GL: {
        // term k of the vector expression = state element k (coefficient x wire, or the coefficient on the ONE wire)
        const auto& off = api.LeOffsets();
        const auto& wi = api.LeWires();
        const auto& ci = api.LeCoeffIds();
        if (off[in.le[0] + 1] - off[in.le[0]] != 12) { g_err = "poseidon-gl macro: input vector must have 12 terms"; return -1; }
        uint64_t st[12];
        for (uint32_t k = 0; k < 12; k++) {
          const uint32_t q = off[in.le[0]] + k;
          const Fr v = wi[q] == 0 ? api.Coeffs()[ci[q]] : mul(api.Coeffs()[ci[q]], c->w[wi[q]]);
          uint64_t x[4];
          canon(v, x);
          if (x[1] | x[2] | x[3] || x[0] >= gl::P) {
            g_err = "MulAddHint: operand is not in the field";  // the permutation starts with gl.Add (base.go:228-232)
            return -5;
          }
          st[k] = x[0];
        }
        std::vector<uint64_t> gt(glm::T_TOTAL);
        memcpy(gt.data() + glm::T_RC, GPW_GL_ALL_ROUND_CONSTANTS, sizeof(GPW_GL_ALL_ROUND_CONSTANTS));
        memcpy(gt.data() + glm::T_CIRC, GPW_GL_MDS_CIRC, sizeof(GPW_GL_MDS_CIRC));
        memcpy(gt.data() + glm::T_DIAG, GPW_GL_MDS_DIAG, sizeof(GPW_GL_MDS_DIAG));
        memcpy(gt.data() + glm::T_FIRST, GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT, sizeof(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT));
        memcpy(gt.data() + glm::T_PRC, GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS));
        memcpy(gt.data() + glm::T_VS, GPW_GL_FAST_PARTIAL_ROUND_VS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_VS));
        memcpy(gt.data() + glm::T_WHATS, GPW_GL_FAST_PARTIAL_ROUND_W_HATS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_W_HATS));
        memcpy(gt.data() + glm::T_INIT, GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX, sizeof(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX));
        if (in.nout != glm::N_OUT) { g_err = "poseidon-gl macro: unexpected output count"; return -1; }
        const auto& mo = api.MacroOuts();
        glm::trace_seq(st, gt.data(), [&](uint32_t slot, const glm::U192& v) {
          const uint64_t l[4] = {v.l[0], v.l[1], v.l[2], 0};
          c->w[mo[in.outs_off + slot]] = fr_from_limbs(l);
        });
        break;
      }
      default: g_err = "unknown opcode"; return -1;
    }
  }
  if (!divs.empty()) {
    std::vector<Fr> den(divs.size()), pref(divs.size());
    Fr run = Fr::one();
    for (size_t i = 0; i < divs.size(); i++) {
      den[i] = eval_le(*c, divs[i]->le[1]);
      pref[i] = run;
      run = mul(run, den[i]);
    }
    Fr irun = inv(run);
    for (size_t i = divs.size(); i-- > 0;) {
      Fr di = mul(irun, pref[i]);
      irun = mul(irun, den[i]);
      c->w[divs[i]->out] = mul(eval_le(*c, divs[i]->le[0]), di);
    }
  }
  return 0;
}

int ct_solve_testdata(void* h, const char* proof_json, const char* vo_json, const uint64_t* x_commit) {
  Circuit* c = (Circuit*)h;
  try {
    gadgets::InputValues iv = gadgets::ParseProofInputs(c->cd, proof_json, vo_json);
    iv.sec.erase(iv.sec.begin(), iv.sec.begin() + c->n_baked);
    return ct_solve_inputs(h, (const uint64_t*)iv.pub.data(), iv.pub.size(), (const uint64_t*)iv.sec.data(), iv.sec.size(), x_commit);
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// number of unsatisfied constraints; first_bad receives the index of the first one (or -1)
uint64_t ct_check(void* h, int64_t* first_bad) {
  Circuit* c = (Circuit*)h;
  const auto& cons = c->api.Constraints();
  uint64_t bad = 0;
  *first_bad = -1;
  for (size_t k = 0; k < cons.size() / 3; k++) {
    Fr l = eval_le(*c, cons[3 * k]), r = eval_le(*c, cons[3 * k + 1]), o = eval_le(*c, cons[3 * k + 2]);
    if (mul(l, r) != o) {
      if (!bad) *first_bad = (int64_t)k;
      bad++;
    }
  }
  return bad;
}

// outputs of every hint of one kind, in tape order; each output as 4 canonical limbs. Returns the number of u64 written.
size_t ct_hint_outputs(void* h, int op, uint64_t* out, size_t cap_u64) {
  Circuit* c = (Circuit*)h;
  size_t n = 0;
  // the builder's log of hint calls (hints fused into a Poseidon macro are not tape instructions of their own)
  for (const auto& hl : c->api.HintLog()) {
    if (hl.first != op) continue;
    if (n + 4 > cap_u64) return n;
    canon(c->w[hl.second], out + n);
    n += 4;
  }
  return n;
}

void ct_wire_values(void* h, uint64_t first, uint64_t count, uint64_t* out) {
  Circuit* c = (Circuit*)h;
  for (uint64_t i = 0; i < count; i++) canon(c->w[first + i], out + 4 * i);
}

// Poseidon-Goldilocks macro (csrc/poseidon_gl_macro.cuh), host build: (1) the specialised 192-bit ReduceHint against the
// general one on n pseudo-random inputs incl. the edges -> number of mismatches; (2) the sequential trace: output state of
// one permutation and the number of slots it emitted.
uint64_t ct_glm_reduce192_mismatches(uint64_t seed, uint64_t n) {
  uint64_t s = seed ? seed : 1, bad = 0;
  auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (uint64_t it = 0; it < n; it++) {
    glm::U192 v{{next(), next(), next() % gl::P}};
    switch (it % 8) {
      case 0: v.l[2] = 0; break;
      case 1: v.l[2] = 0; v.l[1] = 0; break;
      case 2: v.l[2] = gl::P - 1; v.l[1] = ~0ull; v.l[0] = ~0ull; break;
      case 3: v = glm::U192{{gl::P - 1, 0, 0}}; break;
      case 4: v = glm::U192{{gl::P, 0, 0}}; break;
      case 5: v = glm::U192{{0, 0, 0}}; break;
      default: break;
    }
    const uint64_t x[4] = {v.l[0], v.l[1], v.l[2], 0};
    uint64_t q[4], r, q0, q1, r2;
    gl::reduce_hint(x, q, r);
    glm::reduce192(v, q0, q1, r2);
    if (r != r2 || q0 != q[0] || q1 != q[1] || q[2] || q[3]) bad++;
  }
  return bad;
}

uint32_t ct_glm_permute(const uint64_t* in12, uint64_t* out12) {
  std::vector<uint64_t> gt(glm::T_TOTAL);
  memcpy(gt.data() + glm::T_RC, GPW_GL_ALL_ROUND_CONSTANTS, sizeof(GPW_GL_ALL_ROUND_CONSTANTS));
  memcpy(gt.data() + glm::T_CIRC, GPW_GL_MDS_CIRC, sizeof(GPW_GL_MDS_CIRC));
  memcpy(gt.data() + glm::T_DIAG, GPW_GL_MDS_DIAG, sizeof(GPW_GL_MDS_DIAG));
  memcpy(gt.data() + glm::T_FIRST, GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT, sizeof(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT));
  memcpy(gt.data() + glm::T_PRC, GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS));
  memcpy(gt.data() + glm::T_VS, GPW_GL_FAST_PARTIAL_ROUND_VS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_VS));
  memcpy(gt.data() + glm::T_WHATS, GPW_GL_FAST_PARTIAL_ROUND_W_HATS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_W_HATS));
  memcpy(gt.data() + glm::T_INIT, GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX, sizeof(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX));
  uint64_t st[12];
  memcpy(st, in12, sizeof(st));
  std::vector<uint8_t> seen(glm::N_OUT, 0);
  uint32_t emitted = 0;
  glm::trace_seq(st, gt.data(), [&](uint32_t slot, const glm::U192&) {
    if (slot < glm::N_OUT && !seen[slot]) {
      seen[slot] = 1;
      emitted++;
    }
  });
  memcpy(out12, st, sizeof(st));
  return emitted;
}

void ct_wires_mont(void* h, uint64_t* out) {
  Circuit* c = (Circuit*)h;
  memcpy(out, c->w.data(), c->w.size() * sizeof(Fr));
}
}
