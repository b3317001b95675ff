// fri.Chip, plonk.PlonkChip, gates.*, verifier.VerifierChip and the input codec - see gadgets.h.
#include <algorithm>
#include <cmath>
#include <regex>
#include <stdexcept>

#include "gadgets.h"
#include "json.h"
#include "../poseidon_constants.inc"

namespace gpw {
namespace gadgets {

int FriParams::TotalArities() const {
  int r = 0;
  for (auto b : ReductionArityBits) r += (int)b;
  return r;
}

// ---- types: common_circuit_data.json (types/common_data.go:61-127) -----------------------------------------
static FriConfig parse_fri_config(const json::Value& v) {
  FriConfig c;
  c.RateBits = v["rate_bits"].u64();
  c.CapHeight = v["cap_height"].u64();
  c.ProofOfWorkBits = v["proof_of_work_bits"].u64();
  c.NumQueryRounds = v["num_query_rounds"].u64();
  return c;
}

CommonCircuitData ReadCommonCircuitData(const std::string& text) {
  json::Value raw = json::parse(text);
  CommonCircuitData cd;
  const auto& cfg = raw["config"];
  cd.NumWires = cfg["num_wires"].u64();
  cd.NumRoutedWires = cfg["num_routed_wires"].u64();
  cd.NumChallenges = cfg["num_challenges"].u64();
  cd.ConfigFri = parse_fri_config(cfg["fri_config"]);
  const auto& fp = raw["fri_params"];
  cd.Fri.Config = parse_fri_config(fp["config"]);
  cd.Fri.Hiding = fp["hiding"].boolean();
  cd.Fri.DegreeBits = fp["degree_bits"].u64();
  for (const auto& b : fp["reduction_arity_bits"].arr) cd.Fri.ReductionArityBits.push_back(b.u64());
  cd.DegreeBits = cd.Fri.DegreeBits;
  for (const auto& g : raw["gates"].arr) cd.GateIds.push_back(g.str);
  for (const auto& s : raw["selectors_info"]["selector_indices"].arr) cd.SelectorIndices.push_back(s.u64());
  for (const auto& g : raw["selectors_info"]["groups"].arr) cd.SelectorGroups.push_back({g["start"].u64(), g["end"].u64()});
  cd.QuotientDegreeFactor = raw["quotient_degree_factor"].u64();
  cd.NumGateConstraints = raw["num_gate_constraints"].u64();
  cd.NumConstants = raw["num_constants"].u64();
  cd.NumPublicInputs = raw["num_public_inputs"].u64();
  for (const auto& k : raw["k_is"].arr) cd.KIs.push_back(k.u64());
  cd.NumPartialProducts = raw["num_partial_products"].u64();
  if (cd.Fri.Hiding) throw std::runtime_error("Circuit has hiding enabled, which is not supported");  // common_data.go:121-124
  return cd;
}

// ---- input codec -------------------------------------------------------------------------------------------
static std::array<uint64_t, 4> limbs_u64(uint64_t v) { return {v, 0, 0, 0}; }
static std::array<uint64_t, 4> limbs_dec(const std::string& s) {
  // decimal string -> 256-bit LE limbs
  std::array<uint64_t, 4> l = {0, 0, 0, 0};
  for (char ch : s) {
    if (ch < '0' || ch > '9') throw std::runtime_error("bad decimal field element: " + s);
    unsigned __int128 carry = (unsigned)(ch - '0');
    for (int i = 0; i < 4; i++) {
      unsigned __int128 t = (unsigned __int128)l[i] * 10 + carry;
      l[i] = (uint64_t)t;
      carry = t >> 64;
    }
    if (carry) throw std::runtime_error("field element overflows 256 bits");
  }
  return l;
}

InputValues ParseProofInputs(const CommonCircuitData& cd, const std::string& proof_json, const std::string& vo_json) {
  json::Value raw = json::parse(proof_json);
  json::Value vo = json::parse(vo_json);
  InputValues iv;
  auto push_gl = [&](const json::Value& v) { iv.sec.push_back(limbs_u64(v.u64())); };
  auto push_fr = [&](const json::Value& v) { iv.sec.push_back(limbs_dec(v.str)); };
  auto push_qes = [&](const json::Value& arr, size_t expect) {
    if (arr.arr.size() != expect) throw std::runtime_error("proof shape does not match common circuit data");
    for (const auto& q : arr.arr) {
      push_gl(q.arr.at(0));
      push_gl(q.arr.at(1));
    }
  };
  auto push_cap = [&](const json::Value& arr) {
    if (arr.arr.size() != (1ull << cd.Fri.Config.CapHeight)) throw std::runtime_error("cap length mismatch");
    for (const auto& h : arr.arr) push_fr(h);
  };
  const auto& pis = raw["public_inputs"];
  if (pis.arr.size() != cd.NumPublicInputs) throw std::runtime_error("public input count mismatch");
  for (const auto& p : pis.arr) iv.pub.push_back(limbs_u64(p.u64()));
  // verifier-only data
  push_cap(vo["constants_sigmas_cap"]);
  push_fr(vo["circuit_digest"]);
  const auto& p = raw["proof"];
  push_cap(p["wires_cap"]);
  push_cap(p["plonk_zs_partial_products_cap"]);
  push_cap(p["quotient_polys_cap"]);
  const auto& o = p["openings"];
  push_qes(o["constants"], cd.NumConstants);
  push_qes(o["plonk_sigmas"], cd.NumRoutedWires);
  push_qes(o["wires"], cd.NumWires);
  push_qes(o["plonk_zs"], cd.NumChallenges);
  push_qes(o["plonk_zs_next"], cd.NumChallenges);
  push_qes(o["partial_products"], cd.NumChallenges * cd.NumPartialProducts);
  push_qes(o["quotient_polys"], cd.NumChallenges * cd.QuotientDegreeFactor);
  const auto& f = p["opening_proof"];
  if (f["commit_phase_merkle_caps"].arr.size() != cd.Fri.ReductionArityBits.size()) throw std::runtime_error("commit caps mismatch");
  for (const auto& cap : f["commit_phase_merkle_caps"].arr) push_cap(cap);
  const size_t leaf_len[4] = {cd.NumConstants + cd.NumRoutedWires, cd.NumWires, cd.NumChallenges * (1 + cd.NumPartialProducts),
                              cd.NumChallenges * cd.QuotientDegreeFactor};
  if (f["query_round_proofs"].arr.size() != cd.Fri.Config.NumQueryRounds) throw std::runtime_error("query round count mismatch");
  for (const auto& q : f["query_round_proofs"].arr) {
    const auto& eps = q["initial_trees_proof"]["evals_proofs"];
    if (eps.arr.size() != 4) throw std::runtime_error("expected 4 initial trees");
    for (int t = 0; t < 4; t++) {
      const auto& leaf = eps.arr[t].arr.at(0);  // 2-tuple (types/deserialize.go:45-72)
      const auto& sib = eps.arr[t].arr.at(1)["siblings"];
      if (leaf.arr.size() != leaf_len[t]) throw std::runtime_error("leaf length mismatch");
      if (sib.arr.size() != (size_t)(cd.Fri.LdeBits() - (int)cd.Fri.Config.CapHeight)) throw std::runtime_error("sibling count mismatch");
      for (const auto& e : leaf.arr) push_gl(e);
      for (const auto& s : sib.arr) push_fr(s);
    }
    int bits = cd.Fri.LdeBits();
    if (q["steps"].arr.size() != cd.Fri.ReductionArityBits.size()) throw std::runtime_error("steps mismatch");
    for (size_t s = 0; s < cd.Fri.ReductionArityBits.size(); s++) {
      const auto& st = q["steps"].arr[s];
      bits -= (int)cd.Fri.ReductionArityBits[s];
      push_qes(st["evals"], 1ull << cd.Fri.ReductionArityBits[s]);
      const auto& sib = st["merkle_proof"]["siblings"];
      if (sib.arr.size() != (size_t)(bits - (int)cd.Fri.Config.CapHeight)) throw std::runtime_error("step sibling count mismatch");
      for (const auto& x : sib.arr) push_fr(x);
    }
  }
  push_qes(f["final_poly"]["coeffs"], (size_t)cd.Fri.FinalPolyLen());
  push_gl(f["pow_witness"]);
  return iv;
}

// ---- fri (fri/fri_utils.go, fri/fri.go) -----------------------------------------------------------------------
static std::vector<PolynomialInfo> poly_range(uint64_t oracle, uint64_t lo, uint64_t hi) {
  std::vector<PolynomialInfo> v;
  for (uint64_t i = lo; i < hi; i++) v.push_back({oracle, i});
  return v;
}
static uint64_t numPreprocessedPolys(const CommonCircuitData& c) { return c.NumConstants + c.NumRoutedWires; }
static uint64_t numZSPartialProductsPolys(const CommonCircuitData& c) { return c.NumChallenges * (1 + c.NumPartialProducts); }
static uint64_t numQuotientPolys(const CommonCircuitData& c) { return c.NumChallenges * c.QuotientDegreeFactor; }

static void assertNoncanonicalIndicesOK(const FriParams& p) {
  // fri_utils.go:153-163 (config sanity check; the only floating point on the path)
  double num_ambiguous = 18446744073709551615.0 - (double)GL_P + 1.0;
  double query_error = 1.0 / (double)(1ull << p.Config.RateBits);
  if (num_ambiguous / (double)GL_P >= query_error * 1e-5)
    throw std::logic_error("A non-negligible portion of field elements are in the range that permits non-canonical encodings.");
}

FriChip::FriChip(fe::API* api, const CommonCircuitData* cd) : api(api), gl(api), poseidonBN254Chip(api), cd(cd) {}

InstanceInfo FriChip::GetInstance(const QE& zeta) {
  InstanceInfo inst;
  std::vector<PolynomialInfo> all = poly_range(0, 0, numPreprocessedPolys(*cd));
  auto w = poly_range(1, 0, cd->NumWires);
  auto z = poly_range(2, 0, numZSPartialProductsPolys(*cd));
  auto q = poly_range(3, 0, numQuotientPolys(*cd));
  all.insert(all.end(), w.begin(), w.end());
  all.insert(all.end(), z.begin(), z.end());
  all.insert(all.end(), q.begin(), q.end());
  uint64_t g = PrimitiveRootOfUnity(cd->DegreeBits);
  QE zeta_next = gl.MulExtension(gl.CQ(g), zeta);
  inst.Oracles = {{numPreprocessedPolys(*cd), false}, {cd->NumWires, true}, {numZSPartialProductsPolys(*cd), true},
                  {numQuotientPolys(*cd), true}};
  inst.Batches = {{zeta, all}, {zeta_next, poly_range(2, 0, cd->NumChallenges)}};
  return inst;
}

Openings FriChip::ToOpenings(const OpeningSet& c) {
  std::vector<QE> values = c.Constants;
  values.insert(values.end(), c.PlonkSigmas.begin(), c.PlonkSigmas.end());
  values.insert(values.end(), c.Wires.begin(), c.Wires.end());
  values.insert(values.end(), c.PlonkZs.begin(), c.PlonkZs.end());
  values.insert(values.end(), c.PartialProducts.begin(), c.PartialProducts.end());
  values.insert(values.end(), c.QuotientPolys.begin(), c.QuotientPolys.end());
  return {values, c.PlonkZsNext};
}

void FriChip::assertLeadingZeros(const Variable& pow_witness, const FriConfig& cfg) {
  gl.RangeCheckWithMaxBits(pow_witness, 64 - (int)cfg.ProofOfWorkBits);
}

std::vector<QE> FriChip::fromOpeningsAndAlpha(const Openings& openings, const QE& alpha) {
  std::vector<QE> out;
  for (const auto& b : openings) out.push_back(gl.ReduceWithPowers(b, alpha));
  return out;
}

void FriChip::verifyMerkleProofToCapWithCapIndex(const std::vector<Variable>& leaf_data,
                                                 const std::vector<Variable>& leaf_index_bits,
                                                 const std::vector<Variable>& cap_index_bits,
                                                 const std::vector<Variable>& cap, const std::vector<Variable>& siblings) {
  // fri.go:97-144
  Variable current = poseidonBN254Chip.HashOrNoop(leaf_data);
  for (size_t i = 0; i < siblings.size(); i++) {
    const Variable& bit = leaf_index_bits[i];
    Bn254State in = {api->Const(0), api->Const(0), api->Select(bit, siblings[i], current), api->Select(bit, current, siblings[i])};
    current = poseidonBN254Chip.Poseidon(in)[0];
  }
  if (cap_index_bits.size() != 4 || cap.size() != 16)
    throw std::logic_error("capIndexBits length should be 4 and the merkleCap length should be 16");
  Variable leaf_lookups[4];
  for (int i = 0; i < 4; i++)
    leaf_lookups[i] = api->Lookup2(cap_index_bits[0], cap_index_bits[1], cap[4 * i], cap[4 * i + 1], cap[4 * i + 2], cap[4 * i + 3]);
  Variable entry = api->Lookup2(cap_index_bits[2], cap_index_bits[3], leaf_lookups[0], leaf_lookups[1], leaf_lookups[2], leaf_lookups[3]);
  api->AssertIsEqual(current, entry);
}

void FriChip::verifyInitialProof(const std::vector<Variable>& x_index_bits, const std::vector<FriEvalProof>& proofs,
                                 const std::vector<std::vector<Variable>>& caps, const std::vector<Variable>& cap_index_bits) {
  if (proofs.size() != caps.size()) throw std::logic_error("length of eval proofs in fri proof should equal length of initial merkle caps");
  for (size_t i = 0; i < caps.size(); i++)
    verifyMerkleProofToCapWithCapIndex(proofs[i].Elements, x_index_bits, cap_index_bits, caps[i], proofs[i].Siblings);
}

Variable FriChip::expFromBitsConstBase(uint64_t base, const std::vector<Variable>& bits) {
  // fri.go:159-185
  Variable product = gl.C(1);
  for (size_t i = 0; i < bits.size(); i++) {
    uint64_t base_pow = gl_pow(base, 1ull << i);
    Variable base_pow_var = gl.C(base_pow - 1);  // Go: basePow.Uint64() - 1
    product = gl.Add(gl.Mul(gl.Mul(base_pow_var, product), bits[i]), product);
  }
  return product;
}

Variable FriChip::calculateSubgroupX(const std::vector<Variable>& x_index_bits, uint64_t n_log) {
  uint64_t base = PrimitiveRootOfUnity(n_log);
  std::vector<Variable> rev(x_index_bits.rbegin(), x_index_bits.rend());
  Variable product = expFromBitsConstBase(base, rev);
  return gl.Mul(gl.C(7), product);
}

QE FriChip::friCombineInitial(const InstanceInfo& instance, const std::vector<FriEvalProof>& proofs, const QE& alpha,
                              const QE& subgroup_x, const std::vector<QE>& precomputed) {
  // fri.go:208-251
  QE sum = gl.ZeroExtension();
  if (instance.Batches.size() != precomputed.size()) throw std::logic_error("len(openings) != len(precomputedReducedEval)");
  for (size_t i = 0; i < instance.Batches.size(); i++) {
    const auto& batch = instance.Batches[i];
    std::vector<QE> evals;
    for (const auto& p : batch.Polynomials) evals.push_back({proofs[p.OracleIndex].Elements[p.PolynomialInfo_], gl.C(0)});
    QE reduced_evals = gl.ReduceWithPowers(evals, alpha);
    QE numerator = gl.SubExtensionNoReduce(reduced_evals, precomputed[i]);
    QE denominator = gl.SubExtension(subgroup_x, batch.Point);
    sum = gl.MulExtension(gl.ExpExtension(alpha, evals.size()), sum);
    auto inv = gl.InverseExtension(denominator);
    api->AssertIsEqual(inv.second, gl.C(1));
    sum = gl.MulAddExtension(numerator, inv.first, sum);
  }
  return sum;
}

QE FriChip::finalPolyEval(const std::vector<QE>& final_poly, const QE& point) {
  QE ret = gl.ZeroExtension();
  for (size_t k = final_poly.size(); k-- > 0;) ret = gl.MulAddExtension(ret, point, final_poly[k]);
  return ret;
}

QE FriChip::interpolate(const QE& x, const std::vector<QE>& xs, const std::vector<QE>& ys, const std::vector<QE>& weights) {
  // fri.go:261-312
  if (xs.size() != ys.size() || xs.size() != weights.size())
    throw std::logic_error("length of xPoints, yPoints, and barycentricWeights are inconsistent");
  QE l_x = gl.OneExtension();
  for (const auto& xp : xs) l_x = gl.SubMulExtension(x, xp, l_x);
  QE sum = gl.ZeroExtension();
  Variable lookup_from_points = gl.C(1);
  for (size_t i = 0; i < xs.size(); i++) {
    auto q = gl.DivExtension(weights[i], gl.SubExtension(x, xs[i]));
    lookup_from_points = api->Mul(q.second, lookup_from_points);
    sum = gl.AddExtension(gl.MulExtension(ys[i], q.first), sum);
  }
  QE interpolation = gl.MulExtension(l_x, sum);
  QE lookup_val = gl.ZeroExtension();
  for (size_t i = 0; i < xs.size(); i++)
    lookup_val = gl.Lookup(gl.IsZero(gl.SubExtension(x, xs[i])), lookup_val, ys[i]);
  return gl.Lookup(lookup_from_points, lookup_val, interpolation);
}

static uint8_t reverse8(uint8_t b) {
  b = (uint8_t)((b & 0xF0) >> 4 | (b & 0x0F) << 4);
  b = (uint8_t)((b & 0xCC) >> 2 | (b & 0x33) << 2);
  b = (uint8_t)((b & 0xAA) >> 1 | (b & 0x55) << 1);
  return b;
}

QE FriChip::computeEvaluation(const Variable& x, const std::vector<Variable>& within_bits, uint64_t arity_bits,
                              const std::vector<QE>& evals, const QE& beta) {
  // fri.go:314-384
  size_t arity = 1ull << arity_bits;
  if (evals.size() != arity) throw std::logic_error("len(evals) != arity");
  if (arity_bits > 8) throw std::logic_error("currently assuming that arityBits is <= 8");
  uint64_t g = PrimitiveRootOfUnity(arity_bits);
  uint64_t g_inv = gl_pow(g, arity - 1);
  std::vector<QE> permuted(arity);
  for (size_t i = 0; i < arity; i++) permuted[reverse8((uint8_t)i) >> (8 - arity_bits)] = evals[i];
  std::vector<Variable> rev(within_bits.rbegin(), within_bits.rend());
  Variable start = expFromBitsConstBase(g_inv, rev);
  Variable coset_start = gl.Mul(start, x);
  std::vector<QE> xs(arity);
  xs[0] = {coset_start, gl.C(0)};
  for (size_t i = 1; i < arity; i++) xs[i] = gl.MulExtension(xs[i - 1], gl.CQ(g));
  std::vector<QE> weights(arity);
  for (size_t i = 0; i < arity; i++) {
    QE w = gl.OneExtension();
    for (size_t j = 0; j < arity; j++)
      if (i != j) w = gl.SubMulExtension(xs[i], xs[j], w);
    auto inv = gl.InverseExtension(w);
    api->AssertIsEqual(inv.second, gl.C(1));
    weights[i] = inv.first;
  }
  return interpolate(beta, xs, permuted, weights);
}

void FriChip::verifyQueryRound(const InstanceInfo& instance, const FriChallenges& ch, const std::vector<QE>& precomputed,
                               const std::vector<std::vector<Variable>>& caps, const FriProof& proof, Variable x_index,
                               uint64_t n_log, const FriQueryRound& round) {
  // fri.go:386-498
  assertNoncanonicalIndicesOK(cd->Fri);
  x_index = gl.Reduce(x_index);
  std::vector<Variable> all_bits = api->ToBinary(x_index, 64);
  std::vector<Variable> x_index_bits(all_bits.begin(), all_bits.begin() + (cd->Fri.DegreeBits + cd->Fri.Config.RateBits));
  std::vector<Variable> cap_index_bits(x_index_bits.end() - cd->Fri.Config.CapHeight, x_index_bits.end());
  verifyInitialProof(x_index_bits, round.EvalsProofs, caps, cap_index_bits);
  Variable subgroup_x = calculateSubgroupX(x_index_bits, n_log);
  QE old_eval = friCombineInitial(instance, round.EvalsProofs, ch.FriAlpha, {subgroup_x, gl.C(0)}, precomputed);
  for (size_t i = 0; i < cd->Fri.ReductionArityBits.size(); i++) {
    uint64_t arity_bits = cd->Fri.ReductionArityBits[i];
    const std::vector<QE>& evals = round.Steps[i].Evals;
    std::vector<Variable> coset_index_bits(x_index_bits.begin() + arity_bits, x_index_bits.end());
    std::vector<Variable> within(x_index_bits.begin(), x_index_bits.begin() + arity_bits);
    if (arity_bits != 4) throw std::logic_error("assuming arity bits is 4");
    QE leaf[4];
    for (int k = 0; k < 4; k++) leaf[k] = gl.Lookup2(within[0], within[1], evals[4 * k], evals[4 * k + 1], evals[4 * k + 2], evals[4 * k + 3]);
    QE new_eval = gl.Lookup2(within[2], within[3], leaf[0], leaf[1], leaf[2], leaf[3]);
    gl.AssertIsEqual(new_eval[0], old_eval[0]);
    gl.AssertIsEqual(new_eval[1], old_eval[1]);
    old_eval = computeEvaluation(subgroup_x, within, arity_bits, evals, ch.FriBetas[i]);
    std::vector<Variable> field_evals;
    for (const auto& e : evals) {
      field_evals.push_back(e[0]);
      field_evals.push_back(e[1]);
    }
    verifyMerkleProofToCapWithCapIndex(field_evals, coset_index_bits, cap_index_bits, proof.CommitPhaseMerkleCaps[i],
                                       round.Steps[i].Siblings);
    for (uint64_t j = 0; j < arity_bits; j++) subgroup_x = gl.Mul(subgroup_x, subgroup_x);
    x_index_bits = coset_index_bits;
  }
  QE final_eval = finalPolyEval(proof.FinalPoly, {subgroup_x, gl.C(0)});
  gl.AssertIsEqual(old_eval[0], final_eval[0]);
  gl.AssertIsEqual(old_eval[1], final_eval[1]);
}

void FriChip::VerifyFriProof(const InstanceInfo& instance, const Openings& openings, const FriChallenges& ch,
                             const std::vector<std::vector<Variable>>& caps, const FriProof& proof) {
  // fri.go:500-548 (validateFriProofShape is enforced by the input codec, which allocates exactly this shape)
  assertLeadingZeros(ch.FriPowResponse, cd->Fri.Config);
  if (cd->Fri.Config.NumQueryRounds != proof.QueryRoundProofs.size()) throw std::logic_error("Number of query rounds does not match config.");
  std::vector<QE> precomputed = fromOpeningsAndAlpha(openings, ch.FriAlpha);
  uint64_t n_log = cd->Fri.DegreeBits + cd->Fri.Config.RateBits;
  if (ch.FriQueryIndices.size() != proof.QueryRoundProofs.size())
    throw std::logic_error("Number of query indices should equal number of query round proofs");
  for (size_t i = 0; i < ch.FriQueryIndices.size(); i++)
    verifyQueryRound(instance, ch, precomputed, caps, proof, ch.FriQueryIndices[i], n_log, proof.QueryRoundProofs[i]);
}

// ---- gates (plonk/gates/*.go) ------------------------------------------------------------------------------------
namespace {
constexpr uint64_t UNUSED_SELECTOR = 0xffffffffull;

Alg alg_at(const std::vector<QE>& wires, size_t start) { return {wires[start], wires[start + 1]}; }

struct NoopGate : Gate {
  std::string Id() const override { return "NoopGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip*, const EvaluationVars&) override { return {}; }
};
struct ConstantGate : Gate {
  uint64_t n;
  explicit ConstantGate(uint64_t n) : n(n) {}
  std::string Id() const override { return "ConstantGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) c.push_back(g->SubExtension(v.localConstants[i], v.localWires[i]));
    return c;
  }
};
struct PublicInputGate : Gate {
  std::string Id() const override { return "PublicInputGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    std::vector<QE> c;
    for (int i = 0; i < 4; i++) c.push_back(g->SubExtension(v.localWires[i], {v.publicInputsHash[i], g->C(0)}));
    return c;
  }
};
struct ArithmeticGate : Gate {
  uint64_t n;
  explicit ArithmeticGate(uint64_t n) : n(n) {}
  std::string Id() const override { return "ArithmeticGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const QE &c0 = v.localConstants[0], &c1 = v.localConstants[1];
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) {
      const QE &m0 = v.localWires[4 * i], &m1 = v.localWires[4 * i + 1], &ad = v.localWires[4 * i + 2], &out = v.localWires[4 * i + 3];
      // NOTE: C++ leaves the evaluation order of function arguments unspecified; the reference (Go) evaluates left to
      // right and the order of hint creation is part of the witness layout, so nested calls are sequenced explicitly.
      QE lhs = g->MulExtension(g->MulExtension(m0, m1), c0);
      QE rhs = g->MulExtension(ad, c1);
      QE computed = g->AddExtension(lhs, rhs);
      c.push_back(g->SubExtension(out, computed));
    }
    return c;
  }
};
struct ArithmeticExtensionGate : Gate {
  uint64_t n;
  explicit ArithmeticExtensionGate(uint64_t n) : n(n) {}
  std::string Id() const override { return "ArithmeticExtensionGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const QE &c0 = v.localConstants[0], &c1 = v.localConstants[1];
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) {
      Alg m0 = alg_at(v.localWires, 8 * i), m1 = alg_at(v.localWires, 8 * i + 2), ad = alg_at(v.localWires, 8 * i + 4),
          out = alg_at(v.localWires, 8 * i + 6);
      Alg mul = g->MulExtensionAlgebra(m0, m1);
      Alg scaled = g->ScalarMulExtensionAlgebra(c0, mul);
      Alg computed = g->ScalarMulExtensionAlgebra(c1, ad);
      computed = g->AddExtensionAlgebra(computed, scaled);
      Alg diff = g->SubExtensionAlgebra(out, computed);
      c.push_back(diff[0]);
      c.push_back(diff[1]);
    }
    return c;
  }
};
struct MulExtensionGate : Gate {
  uint64_t n;
  explicit MulExtensionGate(uint64_t n) : n(n) {}
  std::string Id() const override { return "MulExtensionGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const QE& c0 = v.localConstants[0];
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) {
      Alg m0 = alg_at(v.localWires, 6 * i), m1 = alg_at(v.localWires, 6 * i + 2), out = alg_at(v.localWires, 6 * i + 4);
      Alg mul = g->MulExtensionAlgebra(m0, m1);
      Alg computed = g->ScalarMulExtensionAlgebra(c0, mul);
      Alg diff = g->SubExtensionAlgebra(out, computed);
      c.push_back(diff[0]);
      c.push_back(diff[1]);
    }
    return c;
  }
};
struct BaseSumGate : Gate {
  uint64_t limbs, base;
  BaseSumGate(uint64_t l, uint64_t b) : limbs(l), base(b) {}
  std::string Id() const override { return "BaseSumGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const QE& sum = v.localWires[0];
    std::vector<QE> ls;
    for (uint64_t i = 0; i < limbs; i++) ls.push_back(v.localWires[1 + i]);
    QE computed = g->ReduceWithPowers(ls, g->CQ(base));
    std::vector<QE> c{g->SubExtension(computed, sum)};
    for (const auto& limb : ls) {
      QE acc = g->OneExtension();
      for (uint64_t i = 0; i < base; i++) acc = g->MulExtension(acc, g->SubExtension(limb, g->CQ(i)));
      c.push_back(acc);
    }
    return c;
  }
};
struct CosetInterpolationGate : Gate {
  uint64_t subgroupBits, degree;
  std::vector<uint64_t> weights;
  CosetInterpolationGate(uint64_t s, uint64_t d, std::vector<uint64_t> w) : subgroupBits(s), degree(d), weights(std::move(w)) {}
  std::string Id() const override { return "CosetInterpolationGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    // coset_interpolation_gate.go:151-226
    const uint64_t n = 1ull << subgroupBits;
    const uint64_t start_values = 1, start_eval_point = start_values + n * 2, start_eval_value = start_eval_point + 2,
                   start_inter = start_eval_value + 2, ni = (n - 2) / (degree - 1);
    std::vector<QE> c;
    const QE& shift = v.localWires[0];
    Alg evaluation_point = alg_at(v.localWires, start_eval_point);
    Alg shifted_point = alg_at(v.localWires, start_inter + 2 * 2 * ni);
    QE neg_shift = g->ScalarMulExtension(shift, g->C(GL_NEG_ONE));
    Alg tmp = g->ScalarMulExtensionAlgebra(neg_shift, shifted_point);
    tmp = g->AddExtensionAlgebra(tmp, evaluation_point);
    c.push_back(tmp[0]);
    c.push_back(tmp[1]);
    std::vector<uint64_t> domain = TwoAdicSubgroup(subgroupBits);
    std::vector<Alg> values;
    for (uint64_t i = 0; i < n; i++) values.push_back(alg_at(v.localWires, start_values + i * 2));
    auto r = g->PartialInterpolateExtAlgebra(domain.data(), values.data(), weights.data(), degree, shifted_point,
                                             g->ZeroExtensionAlgebra(), g->OneExtensionAlgebra());
    Alg c_eval = r.first, c_prod = r.second;
    for (uint64_t i = 0; i < ni; i++) {
      Alg inter_eval = alg_at(v.localWires, start_inter + 2 * i);
      Alg inter_prod = alg_at(v.localWires, start_inter + 2 * (ni + i));
      Alg d = g->SubExtensionAlgebra(inter_eval, c_eval);
      c.push_back(d[0]);
      c.push_back(d[1]);
      d = g->SubExtensionAlgebra(inter_prod, c_prod);
      c.push_back(d[0]);
      c.push_back(d[1]);
      uint64_t s = 1 + (degree - 1) * (i + 1);
      uint64_t e = std::min(s + degree - 1, n);
      auto r2 = g->PartialInterpolateExtAlgebra(domain.data() + s, values.data() + s, weights.data() + s, e - s, shifted_point,
                                                inter_eval, inter_prod);
      c_eval = r2.first;
      c_prod = r2.second;
    }
    Alg evaluation_value = alg_at(v.localWires, start_eval_value);
    Alg d = g->SubExtensionAlgebra(evaluation_value, c_eval);
    c.push_back(d[0]);
    c.push_back(d[1]);
    return c;
  }
};
struct ExponentiationGate : Gate {
  uint64_t n;
  explicit ExponentiationGate(uint64_t n) : n(n) {}
  std::string Id() const override { return "ExponentiationGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const QE& base = v.localWires[0];
    const QE& output = v.localWires[1 + n];
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) {
      QE prev = i == 0 ? g->OneExtension() : g->MulExtension(v.localWires[2 + n + i - 1], v.localWires[2 + n + i - 1]);
      const QE& cur_bit = v.localWires[1 + (n - i - 1)];
      QE tmp = g->MulExtension(cur_bit, g->OneExtension());
      tmp = g->SubExtension(tmp, g->OneExtension());
      QE mul_by = g->MulExtension(cur_bit, base);
      mul_by = g->SubExtension(mul_by, tmp);
      QE diff = g->MulExtension(prev, mul_by);
      diff = g->SubExtension(diff, v.localWires[2 + n + i]);
      c.push_back(diff);
    }
    c.push_back(g->SubExtension(output, v.localWires[2 + n + n - 1]));
    return c;
  }
};
struct PoseidonGate : Gate {
  std::string Id() const override { return "PoseidonGate"; }
  std::vector<QE> EvalUnfiltered(fe::API* api, GlChip* g, const EvaluationVars& v) override {
    // poseidon_gate.go:92-181
    constexpr int W = 12, START_DELTA = 2 * W + 1, START_FULL_0 = START_DELTA + 4, START_PARTIAL = START_FULL_0 + 3 * W,
                  START_FULL_1 = START_PARTIAL + 22;
    PoseidonGlChip pc(api);
    const auto& w = v.localWires;
    std::vector<QE> c;
    const QE& swap = w[2 * W];
    c.push_back(g->MulExtension(swap, g->SubExtension(swap, g->OneExtension())));
    for (int i = 0; i < 4; i++) {
      QE diff = g->SubExtension(w[i + 4], w[i]);
      c.push_back(g->SubExtension(g->MulExtension(swap, diff), w[START_DELTA + i]));
    }
    GlStateExt state;
    for (int i = 0; i < 4; i++) {
      state[i] = g->AddExtension(w[i], w[START_DELTA + i]);
      state[i + 4] = g->SubExtension(w[i + 4], w[START_DELTA + i]);
    }
    for (int i = 8; i < W; i++) state[i] = w[i];
    int rc = 0;
    for (int r = 0; r < 4; r++) {
      state = pc.ConstantLayerExtension(state, &rc);
      if (r != 0)
        for (int i = 0; i < W; i++) {
          const QE& sbox_in = w[START_FULL_0 + (r - 1) * W + i];
          c.push_back(g->SubExtension(state[i], sbox_in));
          state[i] = sbox_in;
        }
      state = pc.SBoxLayerExtension(state);
      state = pc.MdsLayerExtension(state);
      rc++;
    }
    state = pc.PartialFirstConstantLayerExtension(state);
    state = pc.MdsPartialLayerInitExtension(state);
    for (int r = 0; r < 21; r++) {
      const QE& sbox_in = w[START_PARTIAL + r];
      c.push_back(g->SubExtension(state[0], sbox_in));
      state[0] = pc.SBoxMonomialExtension(sbox_in);
      state[0] = g->AddExtension(state[0], g->CQ(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS[r]));
      state = pc.MdsPartialLayerFastExtension(state, r);
    }
    {
      const QE& sbox_in = w[START_PARTIAL + 21];
      c.push_back(g->SubExtension(state[0], sbox_in));
      state[0] = pc.SBoxMonomialExtension(sbox_in);
      state = pc.MdsPartialLayerFastExtension(state, 21);
    }
    rc += 22;
    for (int r = 0; r < 4; r++) {
      state = pc.ConstantLayerExtension(state, &rc);
      for (int i = 0; i < W; i++) {
        const QE& sbox_in = w[START_FULL_1 + r * W + i];
        c.push_back(g->SubExtension(state[i], sbox_in));
        state[i] = sbox_in;
      }
      state = pc.SBoxLayerExtension(state);
      state = pc.MdsLayerExtension(state);
      rc++;
    }
    for (int i = 0; i < W; i++) c.push_back(g->SubExtension(state[i], w[W + i]));
    return c;
  }
};
struct PoseidonMdsGate : Gate {
  std::string Id() const override { return "PoseidonMdsGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    constexpr int W = 12;
    std::vector<Alg> in;
    for (int i = 0; i < W; i++) in.push_back(alg_at(v.localWires, 2 * i));
    std::vector<QE> c;
    std::vector<Alg> computed;
    for (int r = 0; r < W; r++) {
      Alg res = g->ZeroExtensionAlgebra();
      for (int i = 0; i < W; i++)
        res = g->AddExtensionAlgebra(res, g->ScalarMulExtensionAlgebra(g->CQ(GPW_GL_MDS_CIRC[i]), in[(i + r) % W]));
      res = g->AddExtensionAlgebra(res, g->ScalarMulExtensionAlgebra(g->CQ(GPW_GL_MDS_DIAG[r]), in[r]));
      computed.push_back(res);
    }
    for (int i = 0; i < W; i++) {
      Alg out = alg_at(v.localWires, (W + i) * 2);
      Alg diff = g->SubExtensionAlgebra(out, computed[i]);
      c.push_back(diff[0]);
      c.push_back(diff[1]);
    }
    return c;
  }
};
struct RandomAccessGate : Gate {
  uint64_t bits, copies, extra;
  RandomAccessGate(uint64_t b, uint64_t c, uint64_t e) : bits(b), copies(c), extra(e) {}
  std::string Id() const override { return "RandomAccessGate"; }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const uint64_t vec = 1ull << bits, start_extra = (2 + vec) * copies, routed = start_extra + extra;
    const auto& w = v.localWires;
    std::vector<QE> c;
    for (uint64_t copy = 0; copy < copies; copy++) {
      const QE& access_index = w[(2 + vec) * copy];
      std::vector<QE> items;
      for (uint64_t i = 0; i < vec; i++) items.push_back(w[(2 + vec) * copy + 2 + i]);
      const QE& claimed = w[(2 + vec) * copy + 1];
      std::vector<QE> bs;
      for (uint64_t i = 0; i < bits; i++) bs.push_back(w[routed + copy * bits + i]);
      for (const auto& b : bs) c.push_back(g->SubExtension(g->MulExtension(b, b), b));
      QE rec = g->ReduceWithPowers(bs, g->CQ(2));
      c.push_back(g->SubExtension(rec, access_index));
      for (const auto& b : bs) {
        std::vector<QE> next;
        for (size_t i = 0; i < items.size(); i += 2) {
          QE diff = g->SubExtension(items[i + 1], items[i]);
          QE mul = g->MulExtension(b, diff);
          next.push_back(g->AddExtension(items[i], mul));
        }
        items = next;
      }
      if (items.size() != 1) throw std::logic_error("listItems(len) != 1");
      c.push_back(g->SubExtension(items[0], claimed));
    }
    for (uint64_t i = 0; i < extra; i++) c.push_back(g->SubExtension(v.localConstants[i], w[start_extra + i]));
    return c;
  }
};
struct ReducingGateT : Gate {
  uint64_t n;
  bool ext;
  ReducingGateT(uint64_t n, bool ext) : n(n), ext(ext) {}
  std::string Id() const override { return ext ? "ReducingExtensionGate" : "ReducingGate"; }
  uint64_t accs(uint64_t i) const {
    if (i == n - 1) return 0;
    return 6 + (ext ? n * 2 : n) + 2 * i;
  }
  std::vector<QE> EvalUnfiltered(fe::API*, GlChip* g, const EvaluationVars& v) override {
    const auto& w = v.localWires;
    Alg alpha = alg_at(w, 2), acc = alg_at(w, 4);
    std::vector<QE> c;
    for (uint64_t i = 0; i < n; i++) {
      Alg coeff = ext ? alg_at(w, 6 + 2 * i) : Alg{w[6 + i], g->ZeroExtension()};
      Alg acc_i = alg_at(w, accs(i));
      Alg tmp = g->MulExtensionAlgebra(acc, alpha);
      tmp = g->AddExtensionAlgebra(tmp, coeff);
      tmp = g->SubExtensionAlgebra(tmp, acc_i);
      c.push_back(tmp[0]);
      c.push_back(tmp[1]);
      acc = acc_i;
    }
    return c;
  }
};

uint64_t num(const std::smatch& m, int i) { return std::stoull(m[i].str()); }
}  // namespace

std::unique_ptr<Gate> GateInstanceFromId(const std::string& id) {
  // plonk/gates/gates.go:20-54: unanchored, mutually exclusive patterns
  std::smatch m;
  static const std::string PH = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>";
  if (std::regex_search(id, m, std::regex("ArithmeticGate \\{ num_ops: ([0-9]+) \\}"))) return std::make_unique<ArithmeticGate>(num(m, 1));
  if (std::regex_search(id, m, std::regex("ArithmeticExtensionGate \\{ num_ops: ([0-9]+) \\}")))
    return std::make_unique<ArithmeticExtensionGate>(num(m, 1));
  if (std::regex_search(id, m, std::regex("BaseSumGate \\{ num_limbs: ([0-9]+) \\} \\+ Base: ([0-9]+)")))
    return std::make_unique<BaseSumGate>(num(m, 1), num(m, 2));
  if (std::regex_search(id, m, std::regex("ConstantGate \\{ num_consts: ([0-9]+) \\}"))) return std::make_unique<ConstantGate>(num(m, 1));
  if (std::regex_search(id, m,
                        std::regex("CosetInterpolationGate \\{ subgroup_bits: ([0-9]+), degree: ([0-9]+), barycentric_weights: "
                                   "\\[([0-9, ]+)\\], _phantom: " + PH + " \\}<D=2>"))) {
    std::vector<uint64_t> w;
    std::string s = m[3].str(), tok;
    for (size_t i = 0; i <= s.size(); i++) {
      if (i == s.size() || s[i] == ',') {
        size_t a = tok.find_first_not_of(' ');
        if (a != std::string::npos) w.push_back(std::stoull(tok.substr(a)));
        tok.clear();
      } else {
        tok += s[i];
      }
    }
    if (num(m, 2) < 2) throw std::runtime_error("degree must be at least 2 in CosetInterpolationGate");
    return std::make_unique<CosetInterpolationGate>(num(m, 1), num(m, 2), w);
  }
  if (std::regex_search(id, m, std::regex("ExponentiationGate \\{ num_power_bits: ([0-9]+), _phantom: " + PH + " \\}<D=([0-9]+)>")))
    return std::make_unique<ExponentiationGate>(num(m, 1));
  if (std::regex_search(id, m, std::regex("MulExtensionGate \\{ num_ops: ([0-9]+) \\}"))) return std::make_unique<MulExtensionGate>(num(m, 1));
  if (id.find("NoopGate") != std::string::npos) return std::make_unique<NoopGate>();
  if (id.find("PoseidonGate") != std::string::npos) return std::make_unique<PoseidonGate>();
  if (id.find("PoseidonMdsGate") != std::string::npos) return std::make_unique<PoseidonMdsGate>();
  if (id.find("PublicInputGate") != std::string::npos) return std::make_unique<PublicInputGate>();
  if (std::regex_search(id, m,
                        std::regex("RandomAccessGate \\{ bits: ([0-9]+), num_copies: ([0-9]+), num_extra_constants: ([0-9]+), _phantom: " +
                                   PH + " \\}<D=([0-9]+)>")))
    return std::make_unique<RandomAccessGate>(num(m, 1), num(m, 2), num(m, 3));
  if (std::regex_search(id, m, std::regex("ReducingExtensionGate \\{ num_coeffs: ([0-9]+) \\}")))
    return std::make_unique<ReducingGateT>(num(m, 1), true);
  if (std::regex_search(id, m, std::regex("ReducingGate \\{ num_coeffs: ([0-9]+) \\}"))) return std::make_unique<ReducingGateT>(num(m, 1), false);
  throw std::runtime_error("Unknown gate ID " + id);
}

// ---- plonk.PlonkChip (plonk/plonk.go) ----------------------------------------------------------------------------
PlonkChip::PlonkChip(fe::API* api, const CommonCircuitData* cd) : api(api), gl(api), cd(cd) {
  for (const auto& id : cd->GateIds) gates_.push_back(GateInstanceFromId(id));
}

QE PlonkChip::computeFilter(uint64_t row, std::pair<uint64_t, uint64_t> group, const QE& s, bool many) {
  QE product = gl.OneExtension();
  for (uint64_t i = group.first; i < group.second; i++) {
    if (i == row) continue;
    product = gl.MulExtension(product, gl.SubExtension(gl.CQ(i), s));
  }
  if (many) product = gl.MulExtension(product, gl.SubExtension(gl.CQ(UNUSED_SELECTOR), s));
  return product;
}

std::vector<QE> PlonkChip::EvaluateGateConstraints(const EvaluationVars& vars) {
  // evaluate_gates.go:77-105
  std::vector<QE> constraints(cd->NumGateConstraints, gl.ZeroExtension());
  const uint64_t num_selectors = cd->SelectorGroups.size();
  for (size_t i = 0; i < gates_.size(); i++) {
    uint64_t sel = cd->SelectorIndices[i];
    QE filter = computeFilter(i, cd->SelectorGroups[sel], vars.localConstants[sel], num_selectors > 1);
    EvaluationVars v2 = vars;
    v2.localConstants.erase(v2.localConstants.begin(), v2.localConstants.begin() + num_selectors);  // RemovePrefix
    std::vector<QE> unfiltered = gates_[i]->EvalUnfiltered(api, &gl, v2);
    // evalFiltered multiplies ALL constraints by the filter first (evaluate_gates.go:70-73), the accumulation loop
    // follows (evaluate_gates.go:96-101): two passes, so that the hint order matches the reference
    for (auto& u : unfiltered) u = gl.MulExtension(u, filter);
    for (size_t k = 0; k < unfiltered.size(); k++) {
      if (k >= cd->NumGateConstraints) throw std::logic_error("num_constraints() gave too low of a number");
      constraints[k] = gl.AddExtension(constraints[k], unfiltered[k]);
    }
  }
  return constraints;
}

QE PlonkChip::expPowerOf2Extension(QE x) {
  for (uint64_t i = 0; i < cd->DegreeBits; i++) x = gl.MulExtension(x, x);
  return x;
}

QE PlonkChip::evalL0(const QE& x, const QE& x_pow_n) {
  QE eval_zero_poly = gl.SubExtension(x_pow_n, gl.OneExtension());
  QE denominator = gl.SubExtension(gl.ScalarMulExtension(x, gl.C(1ull << cd->DegreeBits)), gl.CQ(1ull << cd->DegreeBits));
  auto q = gl.DivExtension(eval_zero_poly, denominator);
  api->AssertIsEqual(q.second, gl.C(1));
  return q.first;
}

std::vector<QE> PlonkChip::checkPartialProducts(const std::vector<QE>& nums, const std::vector<QE>& dens, uint64_t ch,
                                                const OpeningSet& o) {
  const uint64_t npp = cd->NumPartialProducts, qdf = cd->QuotientDegreeFactor;
  std::vector<QE> accs{o.PlonkZs[ch]};
  accs.insert(accs.end(), o.PartialProducts.begin() + ch * npp, o.PartialProducts.begin() + (ch + 1) * npp);
  accs.push_back(o.PlonkZsNext[ch]);
  std::vector<QE> checks;
  for (uint64_t i = 0; i <= npp; i++) {
    uint64_t s = i * qdf;
    QE nume = nums[s], deno = dens[s];
    for (uint64_t j = 1; j < qdf; j++) {
      nume = gl.MulExtension(nume, nums[s + j]);
      deno = gl.MulExtension(deno, dens[s + j]);
    }
    QE lhs = gl.MulExtension(accs[i], nume);
    QE rhs = gl.MulExtension(accs[i + 1], deno);
    checks.push_back(gl.SubExtension(lhs, rhs));
  }
  return checks;
}

std::vector<QE> PlonkChip::evalVanishingPoly(const EvaluationVars& vars, const ProofChallenges& ch, const OpeningSet& o,
                                             const QE& zeta_pow_n) {
  std::vector<QE> constraint_terms = EvaluateGateConstraints(vars);
  std::vector<QE> s_ids;
  for (uint64_t i = 0; i < cd->NumRoutedWires; i++) s_ids.push_back(gl.ScalarMulExtension(ch.PlonkZeta, gl.C(cd->KIs[i])));
  QE l0_zeta = evalL0(ch.PlonkZeta, zeta_pow_n);
  std::vector<QE> z1_terms, pp_terms;
  for (uint64_t i = 0; i < cd->NumChallenges; i++) {
    z1_terms.push_back(gl.MulExtension(l0_zeta, gl.SubExtension(o.PlonkZs[i], gl.OneExtension())));
    std::vector<QE> nums, dens;
    for (uint64_t j = 0; j < cd->NumRoutedWires; j++) {
      QE wire_plus_gamma = gl.AddExtension(o.Wires[j], {ch.PlonkGammas[i], gl.C(0)});
      nums.push_back(gl.AddExtension(gl.MulExtension({ch.PlonkBetas[i], gl.C(0)}, s_ids[j]), wire_plus_gamma));
      dens.push_back(gl.AddExtension(gl.MulExtension({ch.PlonkBetas[i], gl.C(0)}, o.PlonkSigmas[j]), wire_plus_gamma));
    }
    auto pp = checkPartialProducts(nums, dens, i, o);
    pp_terms.insert(pp_terms.end(), pp.begin(), pp.end());
  }
  std::vector<QE> terms = z1_terms;
  terms.insert(terms.end(), pp_terms.begin(), pp_terms.end());
  terms.insert(terms.end(), constraint_terms.begin(), constraint_terms.end());
  std::vector<QE> reduced(cd->NumChallenges, gl.ZeroExtension());
  for (size_t k = terms.size(); k-- > 0;)
    for (uint64_t j = 0; j < cd->NumChallenges; j++)
      reduced[j] = gl.AddExtension(terms[k], gl.ScalarMulExtension(reduced[j], ch.PlonkAlphas[j]));
  return reduced;
}

void PlonkChip::Verify(const ProofChallenges& ch, const OpeningSet& o, const GlHashOut& pih) {
  QE zeta_pow_n = expPowerOf2Extension(ch.PlonkZeta);
  EvaluationVars vars{o.Constants, o.Wires, pih};
  std::vector<QE> vanishing = evalVanishingPoly(vars, ch, o, zeta_pow_n);
  QE z_h_zeta = gl.SubExtension(zeta_pow_n, gl.OneExtension());
  const uint64_t qdf = cd->QuotientDegreeFactor;
  for (size_t i = 0; i < vanishing.size(); i++) {
    std::vector<QE> chunk(o.QuotientPolys.begin() + i * qdf, o.QuotientPolys.begin() + (i + 1) * qdf);
    QE prod = gl.MulExtension(z_h_zeta, gl.ReduceWithPowers(chunk, zeta_pow_n));
    gl.AssertIsEqualExtension(vanishing[i], prod);
  }
}

// ---- verifier.VerifierChip (verifier/verifier.go) -------------------------------------------------------------------
VerifierChip::VerifierChip(fe::API* api, const CommonCircuitData* cd)
    : api(api), gl(api), poseidonGlChip(api), friChip(api, cd), plonkChip(api, cd), cd(cd) {}

GlHashOut VerifierChip::GetPublicInputsHash(const std::vector<Variable>& pis) { return poseidonGlChip.HashNoPad(pis); }

ProofChallenges VerifierChip::GetChallenges(const Proof& proof, const GlHashOut& pih, const VerifierOnlyCircuitData& vd) {
  ChallengerChip ch(api);
  const uint64_t n = cd->NumChallenges;
  ch.ObserveBN254Hash(vd.CircuitDigest);
  ch.ObserveHash(pih);
  ch.ObserveCap(proof.WiresCap);
  ProofChallenges pc;
  pc.PlonkBetas = ch.GetNChallenges(n);
  pc.PlonkGammas = ch.GetNChallenges(n);
  ch.ObserveCap(proof.PlonkZsPartialProductsCap);
  pc.PlonkAlphas = ch.GetNChallenges(n);
  ch.ObserveCap(proof.QuotientPolysCap);
  pc.PlonkZeta = ch.GetExtensionChallenge();
  ch.ObserveOpenings(friChip.ToOpenings(proof.Openings));
  pc.Fri = ch.GetFriChallenges(proof.OpeningProof.CommitPhaseMerkleCaps, proof.OpeningProof.FinalPoly,
                               proof.OpeningProof.PowWitness, cd->ConfigFri);
  return pc;
}

void VerifierChip::rangeCheckProof(const Proof& proof) {
  const auto& o = proof.Openings;
  for (const auto* grp : {&o.Constants, &o.PlonkSigmas, &o.Wires, &o.PlonkZs, &o.PlonkZsNext, &o.PartialProducts, &o.QuotientPolys})
    for (const auto& q : *grp) gl.RangeCheckQE(q);
  for (const auto& qr : proof.OpeningProof.QueryRoundProofs) {
    for (const auto& ep : qr.EvalsProofs)
      for (const auto& e : ep.Elements) gl.RangeCheck(e);
    for (const auto& st : qr.Steps)
      for (const auto& e : st.Evals) gl.RangeCheckQE(e);
  }
  for (const auto& c : proof.OpeningProof.FinalPoly) gl.RangeCheckQE(c);
  gl.RangeCheck(proof.OpeningProof.PowWitness);
}

void VerifierChip::Verify(const Proof& proof, const std::vector<Variable>& pis, const VerifierOnlyCircuitData& vd) {
  rangeCheckProof(proof);
  GlHashOut pih = GetPublicInputsHash(pis);
  ProofChallenges ch = GetChallenges(proof, pih, vd);
  plonkChip.Verify(ch, proof.Openings, pih);
  std::vector<std::vector<Variable>> caps = {vd.ConstantSigmasCap, proof.WiresCap, proof.PlonkZsPartialProductsCap,
                                             proof.QuotientPolysCap};
  friChip.VerifyFriProof(friChip.GetInstance(ch.PlonkZeta), friChip.ToOpenings(proof.Openings), ch.Fri, caps, proof.OpeningProof);
}

// ---- one gate as a circuit of its own -----------------------------------------------------------------------------------
// spec = "<n_consts>:<n_wires>:<n_constraints>:<gate id as in common_circuit_data.json>". Secret inputs: the local constants,
// the local wires (2 Goldilocks limbs each) and the 4-element public-inputs hash; public inputs: the expected value of every
// constraint (2 limbs each). The circuit asserts Gate::EvalUnfiltered == the public values, so a gate gadget can be checked
// against ANY independent evaluation of the gate's polynomial (tests/test_gate_vectors.py does that for the gates the
// reference has no vectors for: plonk/gates/gates_test.go covers 11 of 14).
void DefineGateCircuit(fe::API* api, const std::string& spec) {
  size_t p1 = spec.find(':'), p2 = spec.find(':', p1 + 1), p3 = spec.find(':', p2 + 1);
  if (p1 == std::string::npos || p2 == std::string::npos || p3 == std::string::npos) throw std::runtime_error("gate circuit spec: n_consts:n_wires:n_constraints:id");
  const size_t n_consts = std::stoul(spec.substr(0, p1)), n_wires = std::stoul(spec.substr(p1 + 1, p2 - p1 - 1)),
               n_out = std::stoul(spec.substr(p2 + 1, p3 - p2 - 1));
  std::unique_ptr<Gate> gate = GateInstanceFromId(spec.substr(p3 + 1));
  std::vector<Variable> expect;
  for (size_t i = 0; i < 2 * n_out; i++) expect.push_back(api->PublicInput());
  EvaluationVars vars;
  for (size_t i = 0; i < n_consts; i++) {
    Variable a = api->SecretInput();
    Variable b = api->SecretInput();
    vars.localConstants.push_back({a, b});
  }
  for (size_t i = 0; i < n_wires; i++) {
    Variable a = api->SecretInput();
    Variable b = api->SecretInput();
    vars.localWires.push_back({a, b});
  }
  for (int i = 0; i < 4; i++) vars.publicInputsHash[i] = api->SecretInput();
  api->EndInputs();
  GlChip gl(api);
  std::vector<QE> out = gate->EvalUnfiltered(api, &gl, vars);
  if (out.size() != n_out) throw std::runtime_error("gate produced " + std::to_string(out.size()) + " constraints, spec says " + std::to_string(n_out));
  for (size_t i = 0; i < n_out; i++) {
    api->AssertIsEqual(gl.Reduce(out[i][0]), expect[2 * i]);
    api->AssertIsEqual(gl.Reduce(out[i][1]), expect[2 * i + 1]);
  }
  api->Finalize();
}

// verifier_only_circuit_data.json alone, in ParseProofInputs order (cap, then digest): the values baked into a bound circuit
std::vector<std::array<uint64_t, 4>> ParseVerifierOnly(const CommonCircuitData& cd, const std::string& vo_json) {
  json::Value vo = json::parse(vo_json);
  std::vector<std::array<uint64_t, 4>> out;
  const auto& cap = vo["constants_sigmas_cap"];
  if (cap.arr.size() != (1ull << cd.Fri.Config.CapHeight)) throw std::runtime_error("cap length mismatch");
  for (const auto& h : cap.arr) out.push_back(limbs_dec(h.str));
  out.push_back(limbs_dec(vo["circuit_digest"].str));
  return out;
}

// ---- ExampleVerifierCircuit (verifier/util.go:10-24) ---------------------------------------------------------------------
// baked == nullptr: proof and verifier-only data are runtime (secret) inputs. Otherwise the first baked->size() values of
// the secret-input order (ParseProofInputs: verifier-only data first, then the proof) are compile-time CONSTANTS, which is
// what the reference's `gnark:"-"` tags do: 2^cap_height + 1 values = VerifierOnlyCircuitData baked (the circuit then only
// accepts proofs of THAT inner circuit); all of them = the reference's ExampleVerifierCircuit as benchmark.go compiles it.
void DefineVerifierCircuit(fe::API* api, const CommonCircuitData& cd, const std::vector<std::array<uint64_t, 4>>* baked) {
  std::vector<Variable> pis;
  for (uint64_t i = 0; i < cd.NumPublicInputs; i++) pis.push_back(api->PublicInput());
  size_t slot = 0;
  auto sec = [&]() {
    const size_t i = slot++;
    if (baked && i < baked->size()) return api->ConstFr(fe::fr_from_limbs((*baked)[i].data()));
    return api->SecretInput();
  };
  auto sec_vec = [&](size_t n) {
    std::vector<Variable> v;
    for (size_t i = 0; i < n; i++) v.push_back(sec());
    return v;
  };
  auto sec_qes = [&](size_t n) {
    std::vector<QE> v;
    for (size_t i = 0; i < n; i++) {
      Variable a = sec();
      Variable b = sec();
      v.push_back({a, b});
    }
    return v;
  };
  const size_t cap = 1ull << cd.Fri.Config.CapHeight;
  VerifierOnlyCircuitData vd;
  vd.ConstantSigmasCap = sec_vec(cap);
  vd.CircuitDigest = sec();
  Proof p;
  p.WiresCap = sec_vec(cap);
  p.PlonkZsPartialProductsCap = sec_vec(cap);
  p.QuotientPolysCap = sec_vec(cap);
  p.Openings.Constants = sec_qes(cd.NumConstants);
  p.Openings.PlonkSigmas = sec_qes(cd.NumRoutedWires);
  p.Openings.Wires = sec_qes(cd.NumWires);
  p.Openings.PlonkZs = sec_qes(cd.NumChallenges);
  p.Openings.PlonkZsNext = sec_qes(cd.NumChallenges);
  p.Openings.PartialProducts = sec_qes(cd.NumChallenges * cd.NumPartialProducts);
  p.Openings.QuotientPolys = sec_qes(cd.NumChallenges * cd.QuotientDegreeFactor);
  for (size_t i = 0; i < cd.Fri.ReductionArityBits.size(); i++) p.OpeningProof.CommitPhaseMerkleCaps.push_back(sec_vec(cap));
  const size_t leaf_len[4] = {cd.NumConstants + cd.NumRoutedWires, cd.NumWires, cd.NumChallenges * (1 + cd.NumPartialProducts),
                              cd.NumChallenges * cd.QuotientDegreeFactor};
  for (uint64_t q = 0; q < cd.Fri.Config.NumQueryRounds; q++) {
    FriQueryRound qr;
    for (int t = 0; t < 4; t++) {
      FriEvalProof ep;
      ep.Elements = sec_vec(leaf_len[t]);
      ep.Siblings = sec_vec(cd.Fri.LdeBits() - cd.Fri.Config.CapHeight);
      qr.EvalsProofs.push_back(ep);
    }
    int bits = cd.Fri.LdeBits();
    for (uint64_t ab : cd.Fri.ReductionArityBits) {
      bits -= (int)ab;
      FriQueryStep st;
      st.Evals = sec_qes(1ull << ab);
      st.Siblings = sec_vec(bits - cd.Fri.Config.CapHeight);
      qr.Steps.push_back(st);
    }
    p.OpeningProof.QueryRoundProofs.push_back(qr);
  }
  p.OpeningProof.FinalPoly = sec_qes(cd.Fri.FinalPolyLen());
  p.OpeningProof.PowWitness = sec();
  api->EndInputs();
  VerifierChip chip(api, &cd);
  chip.Verify(p, pis, vd);
  api->Finalize();
}

}  // namespace gadgets
}  // namespace gpw
