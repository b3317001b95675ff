// Circuit frontend: the subset of gnark's frontend.API the reference's gadgets use, re-designed for a GPU
// solver. Running the gadget code once against this API records
//   (1) an R1CS  (L_k . w) * (R_k . w) = (O_k . w)   - what Groth16 proves, and
//   (2) a solver TAPE: one instruction per wire-defining step (multiplication, hint, inverse, bit /
//       limb decomposition), tagged with its dependency LEVEL, which the CUDA executor replays for a
//       whole batch of proofs (csrc/solver.cu). This replaces gnark's frontend.Compile + the levelled
//       constraint solver (call site benchmark.go:55 and :249).
//
// A Variable is a linear expression over wires with Fr coefficients (constants ride on wire 0 == 1), as in
// gnark's R1CS builder: additions and multiplications by constants are free, a product of two non-constant
// expressions costs one constraint and one internal wire.
//
// Wire numbering: 0 = ONE, then public inputs, then secret inputs, then internal wires in creation order.
#pragma once
#include <cstdint>
#include <cstring>
#include <iosfwd>
#include <map>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../ff.cuh"

namespace gpw {
namespace fe {  // "frontend"

struct Term {
  uint32_t wire;
  Fr coeff;  // Montgomery
};

// Linear expression sum coeff_i * wire_i, terms sorted by wire id, no zero coefficients.
struct Variable {
  std::vector<Term> t;
  bool is_const() const { return t.empty() || (t.size() == 1 && t[0].wire == 0); }
  bool is_zero() const { return t.empty(); }
};

Fr fr_from_u64(uint64_t v);
Fr fr_from_dec(const std::string& s);
Fr fr_from_limbs(const uint64_t l[4]);  // canonical -> Montgomery
void fr_to_limbs(const Fr& a, uint64_t l[4]);  // Montgomery -> canonical

enum Op : uint8_t {
  OP_MUL = 0,          // out = <A> * <B>
  OP_HINT_MULADD = 1,  // (q, r) = divmod(<A> * <B> + <C>, p)      goldilocks/base.go:223  (2 outputs)
  OP_HINT_REDUCE = 2,  // (q, r) = divmod(<A>, p)                   goldilocks/base.go:284  (2 outputs)
  OP_HINT_GLINV = 3,   // out = <A>^-1 mod p (0 -> 0)               goldilocks/base.go:316
  OP_HINT_SPLIT = 4,   // (hi, lo) = (<A> >> 32, <A> & 0xffffffff)  goldilocks/base.go:339  (2 outputs)
  OP_INVZERO = 5,      // out = 1/<A> in Fr, 0 if <A> == 0          gnark api.IsZero's hint
  OP_BITS = 6,         // nout bits of <A>, LSB first               gnark api.ToBinary's hint
  OP_DIV = 7,          // out = <A> / <B> in Fr                     gnark api.DivUnchecked (log-derivative argument)
  OP_DECOMP = 8,       // nout 16-bit limbs of <A>, LSB first       gnark std/rangecheck DecomposeHint
  OP_COUNT = 9,        // out[i] = multiplicity of value i among all limb wires (i < 65536)   CountHint
  OP_COMMIT = 10,      // out = challenge derived from the commitment to the committed wires  api.Commit
  // Macro instruction: one whole Poseidon-BN254 permutation (poseidon/bn254.go:39-45). Inputs: the 4 state
  // expressions; outputs: the x^2, x^4, x^5 wire of every S-box in the order the gadget code creates them (S-boxes
  // of round 0 whose input is a constant are folded by the builder and emit nothing). The R1CS is untouched - the
  // 264 multiplication constraints stay - only the SOLVER computes the permutation natively instead of replaying
  // 264 separate multiplications whose operands are long linear expressions.
  OP_POSEIDON_BN254 = 11,
  // Macro instruction: one whole Poseidon-Goldilocks permutation (poseidon/goldilocks.go:30-37) as the gadget lays it
  // out: 130 MulAddHint + 630 ReduceHint calls and the 472 S-box products, a dependency chain ~264 instructions long
  // (x134 permutations = the solver's whole sequential spine). Input: ONE "vector" expression of 12 terms, term k =
  // state[k] (a wire or a constant; not summed). Outputs: the 1992 wires of those hints / products in the order the
  // gadget creates them - they are NOT consecutive (the range-check wires of every hint sit between them and stay
  // ordinary tape instructions), so the instruction carries an explicit output list (API::MacroOuts). The R1CS and the
  // hint call order are untouched; only the solver evaluates the permutation natively in 64-bit arithmetic.
  OP_POSEIDON_GL = 12,
};

struct Instr {
  uint8_t op;
  uint32_t out;    // first output wire
  uint32_t nout;   // number of consecutive output wires
  uint32_t le[3];  // linear-expression ids of the inputs (NO_LE if unused)
  uint32_t level;
  uint32_t le3 = 0xffffffffu;  // 4th input (macro instructions only)
  uint32_t outs_off = 0xffffffffu;  // scattered outputs: index of the first of `nout` wire ids in API::MacroOuts()
  // wire id of output k
  uint32_t out_wire(uint32_t k, const std::vector<uint32_t>& macro_outs) const {
    return outs_off == 0xffffffffu ? out + k : macro_outs[outs_off + k];
  }
};
constexpr uint32_t NO_LE = 0xffffffffu;

struct HintCounts {
  uint64_t muladd = 0, reduce = 0, glinv = 0, split = 0, invzero = 0, bits = 0, div = 0, decomp = 0, mul = 0;
};

class API {
 public:
  API();
  // ---- inputs -----------------------------------------------------------------------------------
  Variable PublicInput();
  Variable SecretInput();
  void EndInputs() { inputs_closed_ = true; }
  // ---- constants -----------------------------------------------------------------------------------
  Variable Const(uint64_t v) const;
  Variable ConstFr(const Fr& v) const;
  Variable ConstDec(const std::string& dec) const { return ConstFr(fr_from_dec(dec)); }
  // ---- arithmetic (gnark frontend.API) ----------------------------------------------------------------
  Variable Add(const Variable& a, const Variable& b) const;
  Variable Sub(const Variable& a, const Variable& b) const;
  Variable Neg(const Variable& a) const;
  Variable Mul(const Variable& a, const Variable& b);
  Variable MulConst(const Variable& a, const Fr& c) const;
  Variable MulAcc(const Variable& a, const Variable& b, const Variable& c) { return Add(a, Mul(b, c)); }
  Variable IsZero(const Variable& a);
  Variable Select(const Variable& b, const Variable& i1, const Variable& i2);
  Variable Lookup2(const Variable& b0, const Variable& b1, const Variable& i0, const Variable& i1, const Variable& i2,
                   const Variable& i3);
  std::vector<Variable> ToBinary(const Variable& v, int n = 254);
  Variable FromBinary(const std::vector<Variable>& bits, size_t lo, size_t hi) const;
  Variable DivUnchecked(const Variable& a, const Variable& b);
  void AssertIsEqual(const Variable& a, const Variable& b);
  void AssertIsBoolean(const Variable& b);
  // ---- hints ----------------------------------------------------------------------------------------
  // Emits the instruction and returns its output wires. Inputs may be constants (hints are never folded,
  // exactly like gnark's NewHint).
  std::vector<Variable> NewHint(Op op, uint32_t nout, const Variable* a, const Variable* b = nullptr,
                                const Variable* c = nullptr);
  // ---- commit-based range checking (gnark std/rangecheck, frontend.Committer) ---------------------------
  // Replaces the tape instructions created since (tape_begin, wire_begin) - which must all be OP_MULs defining the
  // consecutive wires [wire_begin, NumWires()) - by ONE macro instruction with the given 4 inputs.
  size_t TapeSize() const { return tape_.size(); }
  void FuseAsMacro(Op op, size_t tape_begin, uint32_t wire_begin, const Variable in[4]);
  // Scattered fusion (OP_POSEIDON_GL): between BeginFuse and EndFuse the gadget code marks the instructions it wants
  // computed by the macro (FuseMarkSince(tape size before creating them)); EndFuse removes exactly those from the tape
  // and appends one macro instruction whose outputs are their output wires in creation order. Everything else created
  // in between stays. Returns false (tape untouched) if an input is neither a constant nor a single wire.
  void BeginFuse();
  void FuseMarkSince(size_t tape_before);
  bool EndFuse(Op op, const Variable* in, size_t n_in, uint32_t expect_nout);
  const std::vector<uint32_t>& MacroOuts() const { return macro_outs_; }
  // output wires of every reference hint call (MulAdd / Reduce / Inverse / SplitLimbs) in call order, fused or not
  const std::vector<std::pair<uint8_t, uint32_t>>& HintLog() const { return hint_log_; }
  void RangeCheckCollect(const Variable& v, int bits);  // goldilocks/base.go:411-421, COMMIT_RANGE_CHECKER branch
  // Runs the deferred range-check construction (goldilocks/base.go:423-442 + gnark rangecheck commit):
  // limb decomposition, multiplicity histogram, commitment, log-derivative sums. Call once, at the end.
  void Finalize();
  // Re-levels the tape "as late as possible": an instruction moves to (min level of its consumers) - 1. The
  // long sequential spine of the verifier (challenger sponge -> FRI) keeps its levels, while the ~80 % of
  // instructions that are leaves of the dataflow (range-check splits, IsZero inverses, selects, limb
  // decompositions) sink to a handful of very wide final levels that the GPU runs across all SMs.
  void ScheduleALAP();
  // The schedule the GPU executor uses: the spine (everything of large dependency height) keeps its ASAP levels so
  // that parallel branches stay aligned; the short side branches are gathered into a few wide levels at the end.
  void ScheduleSpineAndTail();

  // ---- compiled circuit -----------------------------------------------------------------------------
  uint32_t NumWires() const { return next_wire_; }
  uint32_t NumPublic() const { return n_public_; }   // excluding ONE
  uint32_t NumSecret() const { return n_secret_; }
  size_t NumConstraints() const { return cons_.size() / 3; }
  const std::vector<Instr>& Tape() const { return tape_; }
  const std::vector<uint32_t>& Constraints() const { return cons_; }  // 3 LE ids per constraint
  const std::vector<uint32_t>& LeOffsets() const { return le_off_; }
  const std::vector<uint32_t>& LeWires() const { return le_wire_; }
  const std::vector<uint32_t>& LeCoeffIds() const { return le_coeff_; }
  const std::vector<Fr>& Coeffs() const { return coeffs_; }
  const HintCounts& Counts() const { return counts_; }
  uint32_t NumLevels() const { return max_level_ + 1; }
  uint32_t CommitLevel() const { return commit_level_; }
  // range-check bookkeeping (valid after Finalize)
  uint32_t LimbWireStart() const { return limb_wire_start_; }
  uint32_t NumLimbWires() const { return n_limb_wires_; }
  uint32_t CountWireStart() const { return count_wire_start_; }
  uint32_t CommitWire() const { return commit_wire_; }
  const std::vector<std::pair<uint32_t, int>>& RangeChecks() const { return rc_; }  // (LE id, bits)
  uint64_t NumRangeCheckedLimbs() const;

  // Compile cache (the reference gave up on r1cs.WriteTo for this circuit, benchmark.go:94-99): everything a COMPILED circuit
  // needs afterwards - R1CS, scheduled tape, macro output lists, hint log, range-check bookkeeping - as one binary blob
  // (native endianness, versioned). The builder-only state (wire levels, interning maps, collected range checks) is not
  // kept: a loaded API is read-only.
  void Serialize(std::ostream& os) const;
  void Deserialize(std::istream& is);  // throws std::runtime_error on a malformed / foreign blob
  bool Scheduled() const { return scheduled_; }

  static constexpr uint32_t COEFF_ONE = 0, COEFF_NEG_ONE = 1;

 private:
  uint32_t new_wire(uint32_t level);
  uint32_t level_of(const Variable& v) const;
  uint32_t intern_le(const Variable& v);
  uint32_t intern_coeff(const Fr& c);
  void add_constraint(const Variable& l, const Variable& r, const Variable& o);
  Variable wire_var(uint32_t w) const;
  bool const_value(const Variable& v, Fr* out) const;

  uint32_t next_wire_ = 1;
  uint32_t n_public_ = 0, n_secret_ = 0;
  bool inputs_closed_ = false;
  bool finalized_ = false;
  bool scheduled_ = false;  // ScheduleSpineAndTail has run (a circuit loaded from the cache is already scheduled)
  std::vector<uint32_t> wire_level_;
  std::vector<uint8_t> wire_bool_;  // known-boolean wires
  std::vector<Instr> tape_;
  std::vector<uint32_t> cons_;
  std::vector<uint32_t> le_off_, le_wire_, le_coeff_;
  std::vector<Fr> coeffs_;
  struct FrHash {
    size_t operator()(const Fr& a) const {
      uint64_t h = 1469598103934665603ull;
      for (int i = 0; i < 8; i++) h = (h ^ a.l[i]) * 1099511628211ull;
      return (size_t)h;
    }
  };
  std::unordered_map<Fr, uint32_t, FrHash> coeff_ids_;
  std::vector<std::pair<uint32_t, int>> rc_;
  HintCounts counts_;
  uint32_t max_level_ = 0;
  uint32_t commit_level_ = 0;
  uint32_t limb_wire_start_ = 0, n_limb_wires_ = 0, count_wire_start_ = 0, commit_wire_ = 0;
  uint32_t le_one_ = NO_LE;
  std::vector<uint32_t> macro_outs_;
  std::vector<std::pair<uint8_t, uint32_t>> hint_log_;  // (op, output wire)
  bool fuse_active_ = false;
  size_t fuse_begin_ = 0;
  std::vector<size_t> fuse_marked_;
  uint32_t intern_raw_le(const std::vector<Term>& terms);
};

}  // namespace fe
}  // namespace gpw
