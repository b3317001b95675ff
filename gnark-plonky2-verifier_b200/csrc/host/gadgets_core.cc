// goldilocks.Chip, Poseidon chips and the challenger - see gadgets.h for the reference mapping.
#include <algorithm>
#include <stdexcept>

#include "gadgets.h"
#include "../poseidon_constants.inc"

namespace gpw {
namespace gadgets {

using fe::Op;

// ---- Goldilocks helpers on plain integers (compile-time constants of the circuit) ------------------------
uint64_t gl_mul(uint64_t a, uint64_t b) {
  unsigned __int128 t = (unsigned __int128)a * b;
  return (uint64_t)(t % GL_P);
}
uint64_t gl_pow(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  while (e) {
    if (e & 1) r = gl_mul(r, a);
    a = gl_mul(a, a);
    e >>= 1;
  }
  return r;
}
uint64_t PrimitiveRootOfUnity(uint64_t n_log) {
  if (n_log > 32) throw std::logic_error("nLog is greater than TWO_ADICITY");
  uint64_t res = GL_POWER_OF_TWO_GENERATOR;
  for (uint64_t i = 0; i < 32 - n_log; i++) res = gl_mul(res, res);
  return res;
}
std::vector<uint64_t> TwoAdicSubgroup(uint64_t n_log) {
  uint64_t g = PrimitiveRootOfUnity(n_log);
  std::vector<uint64_t> res{1};
  for (uint64_t i = 0; i + 1 < (1ull << n_log); i++) res.push_back(gl_mul(res.back(), g));
  return res;
}

// ---- goldilocks.Chip (goldilocks/base.go) -----------------------------------------------------------------
Variable GlChip::MulAdd(const Variable& a, const Variable& b, const Variable& c) {
  // base.go:196-213
  const size_t before = api->TapeSize();
  auto res = api->NewHint(fe::OP_HINT_MULADD, 2, &a, &b, &c);
  api->FuseMarkSince(before);
  const Variable& quotient = res[0];
  const Variable& remainder = res[1];
  Variable lhs = api->MulAcc(c, a, b);
  Variable rhs = api->MulAcc(remainder, C(GL_P), quotient);
  api->AssertIsEqual(lhs, rhs);
  RangeCheck(quotient);
  RangeCheck(remainder);
  return remainder;
}

Variable GlChip::ReduceWithMaxBits(const Variable& x, int max_nb_bits) {
  // base.go:259-281
  const size_t before = api->TapeSize();
  auto res = api->NewHint(fe::OP_HINT_REDUCE, 2, &x);
  api->FuseMarkSince(before);
  const Variable& quotient = res[0];
  api->RangeCheckCollect(quotient, max_nb_bits);
  const Variable& remainder = res[1];
  RangeCheck(remainder);
  api->AssertIsEqual(x, api->Add(api->Mul(quotient, C(GL_P)), remainder));
  return remainder;
}

std::pair<Variable, Variable> GlChip::Inverse(const Variable& x) {
  // base.go:297-313
  Variable inverse = api->NewHint(fe::OP_HINT_GLINV, 1, &x)[0];
  Variable is_zero = api->IsZero(x);
  Variable has_inv = api->Sub(C(1), is_zero);
  RangeCheck(inverse);
  Variable product = Mul(inverse, x);
  Variable to_check = api->Select(has_inv, product, C(1));
  api->AssertIsEqual(to_check, C(1));
  return {inverse, has_inv};
}

void GlChip::RangeCheck(const Variable& x) {
  // base.go:362-400
  auto res = api->NewHint(fe::OP_HINT_SPLIT, 2, &x);
  const Variable& hi = res[0];
  const Variable& lo = res[1];
  api->AssertIsEqual(api->Add(api->Mul(hi, C(1ull << 32)), lo), x);
  api->RangeCheckCollect(hi, 32);
  api->RangeCheckCollect(lo, 32);
  Variable should_check = api->IsZero(api->Sub(hi, C((1ull << 32) - 1)));
  api->AssertIsEqual(api->Select(should_check, lo, C(0)), C(0));
}

QE GlChip::MulExtensionNoReduce(const QE& a, const QE& b) {
  // quadratic_extension.go:65-71
  Variable c0o0 = MulNoReduce(a[0], b[0]);
  Variable c0o1 = MulNoReduce(MulNoReduce(C(GL_W), a[1]), b[1]);
  Variable c0 = AddNoReduce(c0o0, c0o1);
  Variable c1a = MulNoReduce(a[0], b[1]);  // sequenced explicitly: argument evaluation order is unspecified in C++
  Variable c1b = MulNoReduce(a[1], b[0]);
  Variable c1 = AddNoReduce(c1a, c1b);
  return {c0, c1};
}

QE GlChip::InnerProductExtension(const Variable& constant, const QE& starting_acc,
                                 const std::vector<std::array<QE, 2>>& pairs) {
  QE acc = starting_acc;
  for (const auto& p : pairs) {
    QE m = ScalarMulExtension(p[0], constant);
    acc = MulAddExtensionNoReduce(m, p[1], acc);
  }
  return ReduceExtension(acc);
}

std::pair<QE, Variable> GlChip::InverseExtension(const QE& a) {
  // quadratic_extension.go:123-134
  Variable a_is_zero = IsZero(a);
  api->AssertIsEqual(a_is_zero, C(0));
  QE a_pow_r_minus_1 = {a[0], Mul(a[1], C(GL_DTH_ROOT))};
  QE a_pow_r = MulExtension(a_pow_r_minus_1, a);
  auto inv = Inverse(a_pow_r[0]);
  return {ScalarMulExtension(a_pow_r_minus_1, inv.first), inv.second};
}

std::pair<QE, Variable> GlChip::DivExtension(const QE& a, const QE& b) {
  auto bi = InverseExtension(b);
  return {MulExtension(a, bi.first), bi.second};
}

QE GlChip::ExpExtension(const QE& a, uint64_t exponent) {
  if (exponent == 0) return OneExtension();
  if (exponent == 1) return a;
  if (exponent == 2) return MulExtension(a, a);
  QE current = a, product = OneExtension();
  int len = 64 - __builtin_clzll(exponent);
  for (int i = 0; i < len; i++) {
    if (i != 0) current = MulExtension(current, current);
    if ((exponent >> i) & 1) product = MulExtension(product, current);
  }
  return product;
}

QE GlChip::ReduceWithPowers(const std::vector<QE>& terms, const QE& scalar) {
  QE sum = ZeroExtension();
  for (size_t k = terms.size(); k-- > 0;) {
    sum = AddExtensionNoReduce(MulExtensionNoReduce(sum, scalar), terms[k]);
    sum = ReduceExtension(sum);
  }
  return sum;
}

Alg GlChip::MulExtensionAlgebra(const Alg& a, const Alg& b) {
  // quadratic_extension_algebra.go:50-75 with D = 2
  std::vector<std::array<QE, 2>> inner[2], inner_w[2];
  for (int i = 0; i < 2; i++) {
    for (int j = 0; j < 2 - i; j++) inner[(i + j) % 2].push_back({a[i], b[j]});
    for (int j = 2 - i; j < 2; j++) inner_w[(i + j) % 2].push_back({a[i], b[j]});
  }
  Alg product;
  for (int i = 0; i < 2; i++) {
    QE acc = InnerProductExtension(C(GL_W), ZeroExtension(), inner_w[i]);
    product[i] = InnerProductExtension(C(1), acc, inner[i]);
  }
  return product;
}

std::pair<Alg, Alg> GlChip::PartialInterpolateExtAlgebra(const uint64_t* domain, const Alg* values,
                                                         const uint64_t* weights, size_t n, const Alg& point,
                                                         const Alg& initial_eval, const Alg& initial_prod) {
  if (n == 0) throw std::logic_error("Cannot interpolate with no values");
  Alg new_eval = initial_eval, new_prod = initial_prod;
  for (size_t i = 0; i < n; i++) {
    Alg x_alg = {CQ(domain[i]), ZeroExtension()};
    QE weight = CQ(weights[i]);
    Alg term = SubExtensionAlgebra(point, x_alg);
    Alg weighted = ScalarMulExtensionAlgebra(weight, values[i]);
    new_eval = MulExtensionAlgebra(new_eval, term);
    Alg tmp = MulExtensionAlgebra(weighted, new_prod);
    new_eval = AddExtensionAlgebra(new_eval, tmp);
    new_prod = MulExtensionAlgebra(new_prod, term);
  }
  return {new_eval, new_prod};
}

// ---- poseidon.GoldilocksChip (poseidon/goldilocks.go) -------------------------------------------------------
GlState PoseidonGlChip::Poseidon(const GlState& input) {
  num_perms++;
  GlState state = input;
  int rc = 0;
  // same hints, same wires, same constraints; the solver evaluates the whole permutation with one macro instruction
  // (130 MulAdd + 630 Reduce hints and 472 S-box products = 1992 wires, csrc/poseidon_gl_macro.cuh)
  api->BeginFuse();
  state = fullRounds(state, &rc);
  state = partialRounds(state, &rc);
  state = fullRounds(state, &rc);
  api->EndFuse(fe::OP_POSEIDON_GL, input.data(), 12, 1992);
  return state;
}

std::vector<Variable> PoseidonGlChip::HashNToMNoPad(const std::vector<Variable>& input, int nb_outputs) {
  GlState state;
  for (auto& s : state) s = gl.C(0);
  for (size_t i = 0; i < input.size(); i += 8) {
    for (size_t j = 0; j < 8; j++)
      if (i + j < input.size()) state[j] = input[i + j];
    state = Poseidon(state);
  }
  std::vector<Variable> outputs;
  for (;;) {
    for (int i = 0; i < 8; i++) {
      outputs.push_back(state[i]);
      if ((int)outputs.size() == nb_outputs) return outputs;
    }
    state = Poseidon(state);
  }
}

GlHashOut PoseidonGlChip::HashNoPad(const std::vector<Variable>& input) {
  std::vector<Variable> reduced;
  for (const auto& v : input) reduced.push_back(gl.Reduce(v));
  auto out = HashNToMNoPad(reduced, 4);
  return {out[0], out[1], out[2], out[3]};
}

GlState PoseidonGlChip::fullRounds(GlState state, int* rc) {
  for (int i = 0; i < 4; i++) {
    for (int k = 0; k < 12; k++) state[k] = gl.Add(state[k], gl.C(GPW_GL_ALL_ROUND_CONSTANTS[k + 12 * (*rc)]));
    for (int k = 0; k < 12; k++) state[k] = sBoxMonomial(state[k]);
    GlState next;
    for (int r = 0; r < 12; r++) next[r] = mdsRowShf(r, state);
    state = next;
    (*rc)++;
  }
  return state;
}

GlState PoseidonGlChip::partialRounds(GlState state, int* rc) {
  for (int k = 0; k < 12; k++) state[k] = gl.Add(state[k], gl.C(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT[k]));
  state = mdsPartialLayerInit(state);
  for (int i = 0; i < 22; i++) {
    state[0] = sBoxMonomial(state[0]);
    state[0] = gl.Add(state[0], gl.C(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS[i]));
    state = mdsPartialLayerFast(state, i);
  }
  *rc += 22;
  return state;
}

Variable PoseidonGlChip::sBoxMonomial(const Variable& x) {
  // goldilocks.go:138-145
  Variable x2 = gl.MulNoReduce(x, x);
  Variable x3 = gl.MulNoReduce(x, x2);
  x3 = gl.ReduceWithMaxBits(x3, 192);
  Variable x6 = gl.MulNoReduce(x3, x3);
  Variable x7 = gl.MulNoReduce(x, x6);
  return gl.ReduceWithMaxBits(x7, 192);
}

Variable PoseidonGlChip::mdsRowShf(int r, const GlState& v) {
  Variable res = gl.C(0);
  for (int i = 0; i < 12; i++) res = gl.MulAddNoReduce(v[(i + r) % 12], gl.C(GPW_GL_MDS_CIRC[i]), res);
  res = gl.MulAddNoReduce(v[r], gl.C(GPW_GL_MDS_DIAG[r]), res);
  return gl.Reduce(res);
}

GlState PoseidonGlChip::mdsPartialLayerInit(const GlState& state) {
  GlState result;
  for (auto& s : result) s = gl.C(0);
  result[0] = state[0];
  for (int r = 1; r < 12; r++)
    for (int d = 1; d < 12; d++)
      result[d] = gl.MulAddNoReduce(state[r], gl.C(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (d - 1)]), result[d]);
  for (auto& s : result) s = gl.Reduce(s);
  return result;
}

GlState PoseidonGlChip::mdsPartialLayerFast(const GlState& state, int r) {
  Variable d_sum = gl.C(0);
  for (int i = 1; i < 12; i++) d_sum = gl.MulAddNoReduce(state[i], gl.C(GPW_GL_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1]), d_sum);
  Variable d = gl.MulAddNoReduce(state[0], gl.C(GPW_GL_MDS0TO0), d_sum);
  d = gl.Reduce(d);
  GlState result;
  result[0] = d;
  for (int i = 1; i < 12; i++)
    result[i] = gl.MulAddNoReduce(state[0], gl.C(GPW_GL_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1]), state[i]);
  for (auto& s : result) s = gl.Reduce(s);
  return result;
}

GlStateExt PoseidonGlChip::ConstantLayerExtension(GlStateExt state, int* rc) {
  for (int i = 0; i < 12; i++) state[i] = gl.AddExtension(state[i], gl.CQ(GPW_GL_ALL_ROUND_CONSTANTS[i + 12 * (*rc)]));
  return state;
}
QE PoseidonGlChip::SBoxMonomialExtension(const QE& x) {
  QE x2 = gl.MulExtension(x, x);
  QE x4 = gl.MulExtension(x2, x2);
  QE x3 = gl.MulExtension(x, x2);
  return gl.MulExtension(x4, x3);
}
GlStateExt PoseidonGlChip::SBoxLayerExtension(GlStateExt state) {
  for (auto& s : state) s = SBoxMonomialExtension(s);
  return state;
}
QE PoseidonGlChip::MdsRowShfExtension(int r, const GlStateExt& v) {
  QE res = gl.ZeroExtension();
  for (int i = 0; i < 12; i++) {
    QE res1 = gl.MulExtension(v[(i + r) % 12], gl.CQ(GPW_GL_MDS_CIRC[i]));
    res = gl.AddExtension(res, res1);
  }
  res = gl.AddExtension(res, gl.MulExtension(v[r], gl.CQ(GPW_GL_MDS_DIAG[r])));
  return res;
}
GlStateExt PoseidonGlChip::MdsLayerExtension(const GlStateExt& state) {
  GlStateExt result;
  for (int r = 0; r < 12; r++) result[r] = MdsRowShfExtension(r, state);
  return result;
}
GlStateExt PoseidonGlChip::PartialFirstConstantLayerExtension(GlStateExt state) {
  for (int i = 0; i < 12; i++) state[i] = gl.AddExtension(state[i], gl.CQ(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i]));
  return state;
}
GlStateExt PoseidonGlChip::MdsPartialLayerInitExtension(const GlStateExt& state) {
  GlStateExt result;
  for (auto& s : result) s = gl.ZeroExtension();
  result[0] = state[0];
  for (int r = 1; r < 12; r++)
    for (int d = 1; d < 12; d++) {
      QE t = gl.CQ(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX[(r - 1) * 11 + (d - 1)]);
      result[d] = gl.AddExtension(result[d], gl.MulExtension(state[r], t));
    }
  return result;
}
GlStateExt PoseidonGlChip::MdsPartialLayerFastExtension(const GlStateExt& state, int r) {
  QE d = gl.MulExtension(state[0], gl.CQ(GPW_GL_MDS0TO0));
  for (int i = 1; i < 12; i++) {
    QE t = gl.CQ(GPW_GL_FAST_PARTIAL_ROUND_W_HATS[r * 11 + i - 1]);
    d = gl.AddExtension(d, gl.MulExtension(state[i], t));
  }
  GlStateExt result;
  result[0] = d;
  for (int i = 1; i < 12; i++) {
    QE t = gl.CQ(GPW_GL_FAST_PARTIAL_ROUND_VS[r * 11 + i - 1]);
    result[i] = gl.AddExtension(gl.MulExtension(state[0], t), state[i]);
  }
  return result;
}

// ---- poseidon.BN254Chip (poseidon/bn254.go) ---------------------------------------------------------------
static std::vector<Fr> load_fr_table(const uint64_t* mont_limbs, size_t n) {
  std::vector<Fr> v(n);
  for (size_t i = 0; i < n; i++) memcpy(v[i].l, mont_limbs + 4 * i, 32);
  return v;
}

PoseidonBn254Chip::PoseidonBn254Chip(fe::API* api) : api(api) {
  C_ = load_fr_table(GPW_BN_C_MONT, 88);
  S_ = load_fr_table(GPW_BN_S_MONT, 392);
  M_ = load_fr_table(GPW_BN_M_MONT, 16);
  P_ = load_fr_table(GPW_BN_P_MONT, 16);
}

Bn254State PoseidonBn254Chip::Poseidon(Bn254State state) {
  num_perms++;
  const Variable in[4] = {state[0], state[1], state[2], state[3]};
  const size_t tape_begin = api->TapeSize();
  const uint32_t wire_begin = api->NumWires();
  state = ark(state, 0);
  state = fullRounds(state, true);
  state = partialRounds(state);
  state = fullRounds(state, false);
  // same constraints, same wires; the solver computes the whole permutation with one macro instruction
  api->FuseAsMacro(fe::OP_POSEIDON_BN254, tape_begin, wire_begin, in);
  return state;
}

Variable PoseidonBn254Chip::HashNoPad(const std::vector<Variable>& input) {
  Bn254State state = {api->Const(0), api->Const(0), api->Const(0), api->Const(0)};
  // 2^64 and 2^128 as Fr
  Fr two64 = fe::fr_from_u64(1ull << 32);
  two64 = mul(two64, two64);
  Fr pw[3] = {Fr::one(), two64, mul(two64, two64)};
  for (size_t i = 0; i < input.size(); i += 9) {
    size_t end_i = std::min(input.size(), i + 9);
    int state_idx = 0;
    for (size_t j = i; j < end_i; j += 3, state_idx++) {
      size_t end_j = std::min(end_i, j + 3);
      Variable inter = api->Const(0);
      for (size_t k = j; k < end_j; k++) inter = api->Add(inter, api->MulConst(input[k], pw[k - j]));
      state[state_idx + 1] = inter;
    }
    state = Poseidon(state);
  }
  return state[0];
}

Variable PoseidonBn254Chip::HashOrNoop(const std::vector<Variable>& input) {
  if (input.size() <= 3) {
    Fr two64 = fe::fr_from_u64(1ull << 32);
    two64 = mul(two64, two64);
    Fr pw = Fr::one();
    Variable ret = api->Const(0);
    for (const auto& v : input) {
      ret = api->Add(ret, api->MulConst(v, pw));
      pw = mul(pw, two64);
    }
    return ret;
  }
  return HashNoPad(input);
}

std::vector<Variable> PoseidonBn254Chip::ToVec(const Variable& hash) {
  // bn254.go:106-120: 254 bits -> 56-bit chunks
  std::vector<Variable> bits = api->ToBinary(hash, 254);
  std::vector<Variable> out;
  for (size_t i = 0; i < bits.size(); i += 56) out.push_back(api->FromBinary(bits, i, std::min(bits.size(), i + 56)));
  return out;
}

Bn254State PoseidonBn254Chip::fullRounds(Bn254State s, bool is_first) {
  for (int i = 0; i < 3; i++) {
    for (auto& x : s) x = exp5(x);
    s = is_first ? ark(s, (i + 1) * 4) : ark(s, 5 * 4 + 56 + i * 4);
    s = mix(s, M_);
  }
  for (auto& x : s) x = exp5(x);
  if (is_first) {
    s = ark(s, 16);
    s = mix(s, P_);
  } else {
    s = mix(s, M_);
  }
  return s;
}

Bn254State PoseidonBn254Chip::partialRounds(Bn254State s) {
  for (int i = 0; i < 56; i++) {
    s[0] = exp5(s[0]);
    s[0] = api->Add(s[0], api->ConstFr(C_[20 + i]));
    Variable n0 = api->Const(0);
    for (int j = 0; j < 4; j++) n0 = api->Add(n0, api->MulConst(s[j], S_[7 * i + j]));
    for (int k = 1; k < 4; k++) s[k] = api->Add(s[k], api->MulConst(s[0], S_[7 * i + 4 + k - 1]));
    s[0] = n0;
  }
  return s;
}

Bn254State PoseidonBn254Chip::ark(const Bn254State& s, int it) {
  Bn254State r;
  for (int i = 0; i < 4; i++) r[i] = api->Add(s[i], api->ConstFr(C_[it + i]));
  return r;
}

Variable PoseidonBn254Chip::exp5(const Variable& x) {
  Variable x2 = api->Mul(x, x);
  Variable x4 = api->Mul(x2, x2);
  return api->Mul(x4, x);
}

Bn254State PoseidonBn254Chip::mix(const Bn254State& s, const std::vector<Fr>& m) {
  Bn254State r;
  for (int i = 0; i < 4; i++) {
    Variable acc = api->Const(0);
    for (int j = 0; j < 4; j++) acc = api->Add(acc, api->MulConst(s[j], m[j * 4 + i]));
    r[i] = acc;
  }
  return r;
}

// ---- challenger.Chip (challenger/challenger.go) -------------------------------------------------------------
ChallengerChip::ChallengerChip(fe::API* api) : api(api), poseidonChip(api), poseidonBN254Chip(api), gl(api) {
  for (auto& s : spongeState) s = gl.C(0);
}
void ChallengerChip::ObserveElement(const Variable& e) {
  outputBuffer.clear();
  inputBuffer.push_back(e);
  if (inputBuffer.size() == 8) duplexing();
}
void ChallengerChip::ObserveElements(const std::vector<Variable>& es) {
  for (const auto& e : es) ObserveElement(e);
}
void ChallengerChip::ObserveHash(const GlHashOut& h) {
  for (const auto& e : h) ObserveElement(e);
}
void ChallengerChip::ObserveBN254Hash(const Variable& h) { ObserveElements(poseidonBN254Chip.ToVec(h)); }
void ChallengerChip::ObserveCap(const std::vector<Variable>& cap) {
  for (const auto& h : cap) ObserveBN254Hash(h);
}
void ChallengerChip::ObserveExtensionElement(const QE& e) {
  ObserveElement(e[0]);
  ObserveElement(e[1]);
}
void ChallengerChip::ObserveExtensionElements(const std::vector<QE>& es) {
  for (const auto& e : es) ObserveExtensionElement(e);
}
void ChallengerChip::ObserveOpenings(const std::vector<std::vector<QE>>& openings) {
  for (const auto& b : openings) ObserveExtensionElements(b);
}
Variable ChallengerChip::GetChallenge() {
  if (!inputBuffer.empty() || outputBuffer.empty()) duplexing();
  Variable c = outputBuffer.back();  // challenger.go:94-95: LIFO
  outputBuffer.pop_back();
  return c;
}
std::vector<Variable> ChallengerChip::GetNChallenges(uint64_t n) {
  std::vector<Variable> out;
  for (uint64_t i = 0; i < n; i++) out.push_back(GetChallenge());
  return out;
}
QE ChallengerChip::GetExtensionChallenge() {
  auto v = GetNChallenges(2);
  return {v[0], v[1]};
}
GlHashOut ChallengerChip::GetHash() {
  Variable a = GetChallenge(), b = GetChallenge(), c = GetChallenge(), d = GetChallenge();
  return {a, b, c, d};
}
FriChallenges ChallengerChip::GetFriChallenges(const std::vector<std::vector<Variable>>& caps, const std::vector<QE>& final_poly,
                                               const Variable& pow_witness, const FriConfig& config) {
  FriChallenges c;
  c.FriAlpha = GetExtensionChallenge();
  for (const auto& cap : caps) {
    ObserveCap(cap);
    c.FriBetas.push_back(GetExtensionChallenge());
  }
  ObserveExtensionElements(final_poly);
  ObserveElement(pow_witness);
  c.FriPowResponse = GetChallenge();
  c.FriQueryIndices = GetNChallenges(config.NumQueryRounds);
  return c;
}
void ChallengerChip::duplexing() {
  if (inputBuffer.size() > 8) throw std::logic_error("something went wrong");
  num_duplex++;
  for (size_t i = 0; i < inputBuffer.size(); i++) spongeState[i] = gl.Reduce(inputBuffer[i]);
  inputBuffer.clear();
  spongeState = poseidonChip.Poseidon(spongeState);
  outputBuffer.assign(spongeState.begin(), spongeState.begin() + 8);
}

}  // namespace gadgets
}  // namespace gpw
