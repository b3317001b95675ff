// Host-side mirror of the reference's gadget API (same names, argument meaning and error behaviour), written
// against csrc/host/frontend.h instead of gnark's frontend.API:
//   goldilocks.Chip (+ quadratic extension, extension algebra)   goldilocks/*.go
//   poseidon.GoldilocksChip, poseidon.BN254Chip                  poseidon/goldilocks.go, poseidon/bn254.go
//   challenger.Chip                                              challenger/challenger.go
//   fri.Chip                                                     fri/fri.go, fri/fri_utils.go
//   plonk.PlonkChip, gates.*                                     plonk/plonk.go, plonk/gates/*.go
//   verifier.VerifierChip, ExampleVerifierCircuit                verifier/verifier.go, verifier/util.go
//   types / variables (JSON -> shapes / assignments)             types/*.go, variables/*.go
// Running VerifierChip::Verify once against fe::API compiles the verifier circuit (R1CS + solver tape).
#pragma once
#include <array>
#include <memory>
#include <string>
#include <vector>

#include "frontend.h"

namespace gpw {
namespace gadgets {

using fe::Variable;
using GlVar = Variable;                       // goldilocks.Variable{Limb}
using QE = std::array<Variable, 2>;           // goldilocks.QuadraticExtensionVariable
using Alg = std::array<QE, 2>;                // goldilocks.QuadraticExtensionAlgebraVariable

constexpr uint64_t GL_P = 0xffffffff00000001ull;
constexpr uint64_t GL_NEG_ONE = GL_P - 1;
constexpr uint64_t GL_W = 7;
constexpr uint64_t GL_DTH_ROOT = 18446744069414584320ull;
constexpr uint64_t GL_POWER_OF_TWO_GENERATOR = 1753635133440165772ull;
constexpr int RANGE_CHECK_NB_BITS = 144;

uint64_t gl_mul(uint64_t a, uint64_t b);
uint64_t gl_pow(uint64_t a, uint64_t e);
uint64_t PrimitiveRootOfUnity(uint64_t n_log);               // goldilocks/base.go:445
std::vector<uint64_t> TwoAdicSubgroup(uint64_t n_log);       // goldilocks/base.go:456

// ---- goldilocks.Chip -------------------------------------------------------------------------------
class GlChip {
 public:
  explicit GlChip(fe::API* api) : api(api) {}
  fe::API* api;
  Variable C(uint64_t v) const { return api->Const(v); }
  QE CQ(uint64_t v) const { return {api->Const(v), api->Const(0)}; }
  QE ZeroExtension() const { return CQ(0); }
  QE OneExtension() const { return CQ(1); }
  Alg ZeroExtensionAlgebra() const { return {ZeroExtension(), ZeroExtension()}; }
  Alg OneExtensionAlgebra() const { return {OneExtension(), ZeroExtension()}; }

  Variable Add(const Variable& a, const Variable& b) { return MulAdd(a, C(1), b); }
  Variable AddNoReduce(const Variable& a, const Variable& b) { return api->Add(a, b); }
  Variable Sub(const Variable& a, const Variable& b) { return MulAdd(b, C(GL_NEG_ONE), a); }
  Variable SubNoReduce(const Variable& a, const Variable& b) { return api->Add(a, api->Mul(b, C(GL_NEG_ONE))); }
  Variable Mul(const Variable& a, const Variable& b) { return MulAdd(a, b, C(0)); }
  Variable MulNoReduce(const Variable& a, const Variable& b) {
    const size_t before = api->TapeSize();
    Variable v = api->Mul(a, b);
    api->FuseMarkSince(before);  // no-op outside a BeginFuse / EndFuse region (PoseidonGlChip::Poseidon)
    return v;
  }
  Variable MulAdd(const Variable& a, const Variable& b, const Variable& c);
  Variable MulAddNoReduce(const Variable& a, const Variable& b, const Variable& c) { return api->MulAcc(c, a, b); }
  Variable Reduce(const Variable& x) { return ReduceWithMaxBits(x, RANGE_CHECK_NB_BITS); }
  Variable ReduceWithMaxBits(const Variable& x, int max_nb_bits);
  std::pair<Variable, Variable> Inverse(const Variable& x);
  void RangeCheck(const Variable& x);
  void RangeCheckWithMaxBits(const Variable& x, int bits) { api->RangeCheckCollect(x, bits); }
  void AssertIsEqual(const Variable& x, const Variable& y) { api->AssertIsEqual(x, y); }

  QE AddExtension(const QE& a, const QE& b) { return {Add(a[0], b[0]), Add(a[1], b[1])}; }
  QE AddExtensionNoReduce(const QE& a, const QE& b) { return {AddNoReduce(a[0], b[0]), AddNoReduce(a[1], b[1])}; }
  QE SubExtension(const QE& a, const QE& b) { return {Sub(a[0], b[0]), Sub(a[1], b[1])}; }
  QE SubExtensionNoReduce(const QE& a, const QE& b) { return {SubNoReduce(a[0], b[0]), SubNoReduce(a[1], b[1])}; }
  QE MulExtension(const QE& a, const QE& b) { return ReduceExtension(MulExtensionNoReduce(a, b)); }
  QE MulExtensionNoReduce(const QE& a, const QE& b);
  QE MulAddExtension(const QE& a, const QE& b, const QE& c) {
    return ReduceExtension(AddExtensionNoReduce(MulExtensionNoReduce(a, b), c));
  }
  QE MulAddExtensionNoReduce(const QE& a, const QE& b, const QE& c) {
    return AddExtensionNoReduce(MulExtensionNoReduce(a, b), c);
  }
  QE SubMulExtension(const QE& a, const QE& b, const QE& c) {
    return ReduceExtension(MulExtensionNoReduce(SubExtensionNoReduce(a, b), c));
  }
  QE ScalarMulExtension(const QE& a, const Variable& b) { return {Mul(a[0], b), Mul(a[1], b)}; }
  QE InnerProductExtension(const Variable& constant, const QE& starting_acc, const std::vector<std::array<QE, 2>>& pairs);
  std::pair<QE, Variable> InverseExtension(const QE& a);
  std::pair<QE, Variable> DivExtension(const QE& a, const QE& b);
  QE ExpExtension(const QE& a, uint64_t exponent);
  QE ReduceExtension(const QE& x) { return {Reduce(x[0]), Reduce(x[1])}; }
  QE ReduceWithPowers(const std::vector<QE>& terms, const QE& scalar);
  Variable IsZero(const QE& x) {
    Variable z0 = api->IsZero(x[0]);
    Variable z1 = api->IsZero(x[1]);
    return api->Mul(z0, z1);
  }
  QE Lookup(const Variable& b, const QE& x, const QE& y) {
    return {api->Select(b, y[0], x[0]), api->Select(b, y[1], x[1])};
  }
  QE Lookup2(const Variable& b0, const Variable& b1, const QE& q0, const QE& q1, const QE& q2, const QE& q3) {
    QE c0 = Lookup(b0, q0, q1), c1 = Lookup(b0, q2, q3);
    return Lookup(b1, c0, c1);
  }
  void AssertIsEqualExtension(const QE& a, const QE& b) {
    AssertIsEqual(a[0], b[0]);
    AssertIsEqual(a[1], b[1]);
  }
  void RangeCheckQE(const QE& a) {
    RangeCheck(a[0]);
    RangeCheck(a[1]);
  }

  Alg AddExtensionAlgebra(const Alg& a, const Alg& b) { return {AddExtension(a[0], b[0]), AddExtension(a[1], b[1])}; }
  Alg SubExtensionAlgebra(const Alg& a, const Alg& b) { return {SubExtension(a[0], b[0]), SubExtension(a[1], b[1])}; }
  Alg MulExtensionAlgebra(const Alg& a, const Alg& b);
  Alg ScalarMulExtensionAlgebra(const QE& a, const Alg& b) { return {MulExtension(a, b[0]), MulExtension(a, b[1])}; }
  std::pair<Alg, Alg> PartialInterpolateExtAlgebra(const uint64_t* domain, const Alg* values, const uint64_t* weights,
                                                   size_t n, const Alg& point, const Alg& initial_eval,
                                                   const Alg& initial_prod);
};

// ---- poseidon ------------------------------------------------------------------------------------------
using GlState = std::array<Variable, 12>;
using GlStateExt = std::array<QE, 12>;
using GlHashOut = std::array<Variable, 4>;
using Bn254State = std::array<Variable, 4>;

class PoseidonGlChip {
 public:
  explicit PoseidonGlChip(fe::API* api) : api(api), gl(api) {}
  fe::API* api;
  GlChip gl;
  uint64_t num_perms = 0;
  GlState Poseidon(const GlState& input);
  std::vector<Variable> HashNToMNoPad(const std::vector<Variable>& input, int nb_outputs);
  GlHashOut HashNoPad(const std::vector<Variable>& input);
  // extension-field layers (PoseidonGate)
  GlStateExt ConstantLayerExtension(GlStateExt state, int* rc);
  QE SBoxMonomialExtension(const QE& x);
  GlStateExt SBoxLayerExtension(GlStateExt state);
  GlStateExt MdsLayerExtension(const GlStateExt& state);
  GlStateExt PartialFirstConstantLayerExtension(GlStateExt state);
  GlStateExt MdsPartialLayerInitExtension(const GlStateExt& state);
  GlStateExt MdsPartialLayerFastExtension(const GlStateExt& state, int r);

 private:
  GlState fullRounds(GlState state, int* rc);
  GlState partialRounds(GlState state, int* rc);
  Variable sBoxMonomial(const Variable& x);
  Variable mdsRowShf(int r, const GlState& v);
  GlState mdsPartialLayerInit(const GlState& state);
  GlState mdsPartialLayerFast(const GlState& state, int r);
  QE MdsRowShfExtension(int r, const GlStateExt& v);
};

class PoseidonBn254Chip {
 public:
  explicit PoseidonBn254Chip(fe::API* api);
  fe::API* api;
  uint64_t num_perms = 0;
  Bn254State Poseidon(Bn254State state);
  Variable HashNoPad(const std::vector<Variable>& input);
  Variable HashOrNoop(const std::vector<Variable>& input);
  Variable TwoToOne(const Variable& l, const Variable& r) { return Poseidon({api->Const(0), api->Const(0), l, r})[0]; }
  std::vector<Variable> ToVec(const Variable& hash);

 private:
  Bn254State fullRounds(Bn254State s, bool is_first);
  Bn254State partialRounds(Bn254State s);
  Bn254State ark(const Bn254State& s, int it);
  Variable exp5(const Variable& x);
  Bn254State mix(const Bn254State& s, const std::vector<Fr>& m);
  std::vector<Fr> C_, S_, M_, P_;
};

// ---- types / variables ---------------------------------------------------------------------------------
struct FriConfig {
  uint64_t RateBits = 0, CapHeight = 0, ProofOfWorkBits = 0, NumQueryRounds = 0;
};
struct FriParams {
  FriConfig Config;
  bool Hiding = false;
  uint64_t DegreeBits = 0;
  std::vector<uint64_t> ReductionArityBits;
  int TotalArities() const;
  int LdeBits() const { return (int)(DegreeBits + Config.RateBits); }
  int FinalPolyLen() const { return 1 << ((int)DegreeBits - TotalArities()); }
};
struct CommonCircuitData {
  uint64_t NumWires = 0, NumRoutedWires = 0, NumChallenges = 0;
  FriConfig ConfigFri;
  FriParams Fri;
  uint64_t DegreeBits = 0;
  std::vector<std::string> GateIds;
  std::vector<uint64_t> SelectorIndices;
  std::vector<std::pair<uint64_t, uint64_t>> SelectorGroups;
  uint64_t QuotientDegreeFactor = 0, NumGateConstraints = 0, NumConstants = 0, NumPublicInputs = 0, NumPartialProducts = 0;
  std::vector<uint64_t> KIs;
};
CommonCircuitData ReadCommonCircuitData(const std::string& json_text);  // types/common_data.go:61

// Shape of a proof (variables.New* allocators, variables/fri.go:12-66) + flat input ordering
struct OpeningSet {
  std::vector<QE> Constants, PlonkSigmas, Wires, PlonkZs, PlonkZsNext, PartialProducts, QuotientPolys;
};
struct FriEvalProof {
  std::vector<Variable> Elements;
  std::vector<Variable> Siblings;
};
struct FriQueryStep {
  std::vector<QE> Evals;
  std::vector<Variable> Siblings;
};
struct FriQueryRound {
  std::vector<FriEvalProof> EvalsProofs;
  std::vector<FriQueryStep> Steps;
};
struct FriProof {
  std::vector<std::vector<Variable>> CommitPhaseMerkleCaps;
  std::vector<FriQueryRound> QueryRoundProofs;
  std::vector<QE> FinalPoly;
  Variable PowWitness;
};
struct Proof {
  std::vector<Variable> WiresCap, PlonkZsPartialProductsCap, QuotientPolysCap;
  OpeningSet Openings;
  FriProof OpeningProof;
};
struct VerifierOnlyCircuitData {
  std::vector<Variable> ConstantSigmasCap;
  Variable CircuitDigest;
};

// Canonical flattening of (public inputs, proof, verifier-only data) into input-wire order. The same walk is
// used (a) to allocate circuit inputs and (b) to turn the three JSON files into the input-value vector.
struct InputValues {
  std::vector<std::array<uint64_t, 4>> pub;  // canonical limbs
  std::vector<std::array<uint64_t, 4>> sec;
};
InputValues ParseProofInputs(const CommonCircuitData& cd, const std::string& proof_json, const std::string& verifier_only_json);

struct FriChallenges {
  QE FriAlpha;
  std::vector<QE> FriBetas;
  Variable FriPowResponse;
  std::vector<Variable> FriQueryIndices;
};
struct ProofChallenges {
  std::vector<Variable> PlonkBetas, PlonkGammas, PlonkAlphas;
  QE PlonkZeta;
  FriChallenges Fri;
};

// ---- challenger ----------------------------------------------------------------------------------------
class ChallengerChip {
 public:
  explicit ChallengerChip(fe::API* api);
  void ObserveElement(const Variable& e);
  void ObserveElements(const std::vector<Variable>& es);
  void ObserveHash(const GlHashOut& h);
  void ObserveBN254Hash(const Variable& h);
  void ObserveCap(const std::vector<Variable>& cap);
  void ObserveExtensionElement(const QE& e);
  void ObserveExtensionElements(const std::vector<QE>& es);
  void ObserveOpenings(const std::vector<std::vector<QE>>& openings);
  Variable GetChallenge();
  std::vector<Variable> GetNChallenges(uint64_t n);
  QE GetExtensionChallenge();
  GlHashOut GetHash();
  FriChallenges GetFriChallenges(const std::vector<std::vector<Variable>>& caps, const std::vector<QE>& final_poly,
                                 const Variable& pow_witness, const FriConfig& config);
  uint64_t num_duplex = 0;

 private:
  void duplexing();
  fe::API* api;
  PoseidonGlChip poseidonChip;
  PoseidonBn254Chip poseidonBN254Chip;
  GlChip gl;
  GlState spongeState;
  std::vector<Variable> inputBuffer, outputBuffer;
};

// ---- fri -----------------------------------------------------------------------------------------------
struct PolynomialInfo {
  uint64_t OracleIndex, PolynomialInfo_;
};
struct BatchInfo {
  QE Point;
  std::vector<PolynomialInfo> Polynomials;
};
struct OracleInfo {
  uint64_t NumPolys;
  bool Blinding;
};
struct InstanceInfo {
  std::vector<OracleInfo> Oracles;
  std::vector<BatchInfo> Batches;
};
using Openings = std::vector<std::vector<QE>>;  // fri.Openings{Batches[].Values}

class FriChip {
 public:
  FriChip(fe::API* api, const CommonCircuitData* cd);
  InstanceInfo GetInstance(const QE& zeta);
  Openings ToOpenings(const OpeningSet& c);
  void VerifyFriProof(const InstanceInfo& instance, const Openings& openings, const FriChallenges& ch,
                      const std::vector<std::vector<Variable>>& initial_merkle_caps, const FriProof& proof);

 private:
  void assertLeadingZeros(const Variable& pow_witness, const FriConfig& cfg);
  std::vector<QE> fromOpeningsAndAlpha(const Openings& openings, const QE& alpha);
  void verifyMerkleProofToCapWithCapIndex(const std::vector<Variable>& leaf_data, const std::vector<Variable>& leaf_index_bits,
                                          const std::vector<Variable>& cap_index_bits, const std::vector<Variable>& cap,
                                          const std::vector<Variable>& siblings);
  void verifyInitialProof(const std::vector<Variable>& x_index_bits, const std::vector<FriEvalProof>& proofs,
                          const std::vector<std::vector<Variable>>& caps, const std::vector<Variable>& cap_index_bits);
  Variable expFromBitsConstBase(uint64_t base, const std::vector<Variable>& bits);
  Variable calculateSubgroupX(const std::vector<Variable>& x_index_bits, uint64_t n_log);
  QE friCombineInitial(const InstanceInfo& instance, const std::vector<FriEvalProof>& proofs, const QE& alpha,
                       const QE& subgroup_x, const std::vector<QE>& precomputed);
  QE finalPolyEval(const std::vector<QE>& final_poly, const QE& point);
  QE interpolate(const QE& x, const std::vector<QE>& xs, const std::vector<QE>& ys, const std::vector<QE>& weights);
  QE computeEvaluation(const Variable& x, const std::vector<Variable>& within_bits, uint64_t arity_bits,
                       const std::vector<QE>& evals, const QE& beta);
  void verifyQueryRound(const InstanceInfo& instance, const FriChallenges& ch, const std::vector<QE>& precomputed,
                        const std::vector<std::vector<Variable>>& caps, const FriProof& proof, Variable x_index,
                        uint64_t n_log, const FriQueryRound& round);
  fe::API* api;
  GlChip gl;
  PoseidonBn254Chip poseidonBN254Chip;
  const CommonCircuitData* cd;
};

// ---- plonk / gates -------------------------------------------------------------------------------------
struct EvaluationVars {
  std::vector<QE> localConstants, localWires;
  GlHashOut publicInputsHash;
};
class Gate {
 public:
  virtual ~Gate() {}
  virtual std::string Id() const = 0;
  virtual std::vector<QE> EvalUnfiltered(fe::API* api, GlChip* gl, const EvaluationVars& vars) = 0;
};
std::unique_ptr<Gate> GateInstanceFromId(const std::string& gate_id);  // plonk/gates/gates.go:37

class PlonkChip {
 public:
  PlonkChip(fe::API* api, const CommonCircuitData* cd);
  void Verify(const ProofChallenges& ch, const OpeningSet& openings, const GlHashOut& pih);
  std::vector<QE> EvaluateGateConstraints(const EvaluationVars& vars);

 private:
  QE expPowerOf2Extension(QE x);
  QE evalL0(const QE& x, const QE& x_pow_n);
  std::vector<QE> checkPartialProducts(const std::vector<QE>& nums, const std::vector<QE>& dens, uint64_t challenge_num,
                                       const OpeningSet& openings);
  std::vector<QE> evalVanishingPoly(const EvaluationVars& vars, const ProofChallenges& ch, const OpeningSet& openings,
                                    const QE& zeta_pow_n);
  QE computeFilter(uint64_t row, std::pair<uint64_t, uint64_t> group, const QE& s, bool many);
  fe::API* api;
  GlChip gl;
  const CommonCircuitData* cd;
  std::vector<std::unique_ptr<Gate>> gates_;
};

// ---- verifier ------------------------------------------------------------------------------------------
class VerifierChip {
 public:
  VerifierChip(fe::API* api, const CommonCircuitData* cd);
  GlHashOut GetPublicInputsHash(const std::vector<Variable>& public_inputs);
  ProofChallenges GetChallenges(const Proof& proof, const GlHashOut& pih, const VerifierOnlyCircuitData& vd);
  void Verify(const Proof& proof, const std::vector<Variable>& public_inputs, const VerifierOnlyCircuitData& vd);

 private:
  void rangeCheckProof(const Proof& proof);
  fe::API* api;
  GlChip gl;
  PoseidonGlChip poseidonGlChip;
  FriChip friChip;
  PlonkChip plonkChip;
  const CommonCircuitData* cd;
};

// ExampleVerifierCircuit.Define (verifier/util.go:19-24) with the proof and verifier-only data as SECRET inputs
// (the reference's own test circuits use this form, fri/fri_test.go:17-21) and PublicInputs as public inputs.
// A single gate's EvalUnfiltered as a circuit (see the definition for the spec string and the input order).
void DefineGateCircuit(fe::API* api, const std::string& spec);
std::vector<std::array<uint64_t, 4>> ParseVerifierOnly(const CommonCircuitData& cd, const std::string& verifier_only_json);
// Allocates the inputs in ParseProofInputs order, runs Verify and Finalize. `baked`: leading secret-input slots that
// become compile-time constants instead (see the definition).
void DefineVerifierCircuit(fe::API* api, const CommonCircuitData& cd, const std::vector<std::array<uint64_t, 4>>* baked = nullptr);

}  // namespace gadgets
}  // namespace gpw
