// Sparse constraint system (PLONK gates) derived from a compiled circuit - the counterpart of gnark's scs.NewBuilder
// (frontend/cs/scs), which benchmark.go:44-45 selects for `-proof-system plonk`, and the input of plonk.Setup / plonk.Prove
// (benchmark.go:130, 162).
//
// gnark's scs builder runs the gadget code a second time and emits one gate per api call. Here the SAME compiled circuit
// (R1CS + solver tape, host/frontend.h) is lowered instead, so the witness solver, its hints and its tests stay shared between
// the two backends:
//   gate:   qL a + qR b + qM a b + qO c + qC + PI + Qcp P2 = 0        (one row of the evaluation domain)
//   * every linear expression with two or more terms becomes a chain of addition gates over fresh variables
//       v_1 = c_0 w_0 + c_1 w_1,  v_j = v_(j-1) + c_j w_j
//     (long expressions - the log-derivative sums have millions of terms - are cut into chunks of SCS_CHUNK terms whose sums
//     are chained again, so that no chain is longer than SCS_CHUNK and all chains of a level are independent);
//   * every R1CS row (L, R, O) becomes one multiplication gate over the (scaled) variables its sides reduce to;
//   * public inputs (ONE, the public wires, and - as in Groth16 - the range-check commitment challenge) get one row each:
//       a - x_i = 0, the x_i entering through the public-input polynomial PI;
//   * the committed wires of the range-check argument (limbs + multiplicities) get one row each with Qcp = 1:
//       -a + P2 = 0, which makes the prover-committed polynomial P2 carry exactly those wire values: [P2] is fixed before
//     the challenge is derived from it (the BSB22 scheme gnark's PLONK backend uses for api.Commit).
// Copy constraints: three columns x N rows of variable ids -> permutation sigma (cycles over the slots of a variable).
#pragma once
#include <cstdint>
#include <vector>

#include "frontend.h"

namespace gpw {
namespace scs {

constexpr uint32_t SCS_CHUNK = 256;

// v[out] = konst + c[0] v[w[0]] + c[1] v[w[1]],  v[out + j] = v[out + j - 1] + c[j + 1] v[w[j + 1]]   (j < n_terms - 1)
// konst: coefficient id of the expression's constant term (C_ZERO if none) - it rides on the first gate's qC, which saves the
// gate a term on the ONE wire would cost. n_terms == 1 (a constant plus one scaled wire): v[out] = konst + c[0] v[w[0]].
struct Chain {
  uint32_t term_off;  // into chain_wire / chain_coeff
  uint32_t n_terms;   // >= 1
  uint32_t out;       // first of the max(n_terms - 1, 1) new variables
  uint32_t konst;
};

struct System {
  uint32_t n_orig = 0;      // variables [0, n_orig) are the circuit's wires
  uint32_t n_vars = 0;      // + the chain variables
  uint32_t n_public_rows = 0;  // rows [0, n_public_rows): ONE, public wires, commitment challenge (if any)
  uint32_t n_qcp_rows = 0;     // rows [n_public_rows, n_public_rows + n_qcp_rows): committed wires
  uint32_t n_gates = 0;        // all rows in use; the domain is the next power of two
  int logN = 0;
  bool has_commit = false;
  uint32_t commit_wire = 0, committed_lo = 0, n_committed = 0;
  // per gate (row)
  std::vector<uint32_t> a, b, c;             // variable ids (unused slots hold variable 0)
  std::vector<uint32_t> ql, qr, qm, qo, qc;  // ids into `coeffs`
  std::vector<uint8_t> qcp;
  std::vector<Fr> coeffs;                    // Montgomery; [0] = 0, [1] = 1, [2] = -1
  // witness extension program, level by level (chains of level l read variables of levels < l only)
  std::vector<Chain> chains;
  std::vector<uint32_t> level_off;           // chains [level_off[l], level_off[l + 1])
  std::vector<uint32_t> chain_wire, chain_coeff;
  // public row i constrains variable public_var[i]
  std::vector<uint32_t> public_var;
  static constexpr uint32_t C_ZERO = 0, C_ONE = 1, C_NEG_ONE = 2;
};

// Lowers the compiled circuit. Throws std::runtime_error if the result does not fit 2^27 rows.
System Build(const fe::API& api);

// sigma over the 3 N slots (slot = column * N + row): the next slot of the same variable, cyclically.
void BuildPermutation(const System& s, std::vector<uint32_t>* sigma);

// Host reference of the witness extension and of the gate equations (tests): v has n_vars entries, the first n_orig filled.
void ExtendWitness(const System& s, std::vector<Fr>* v);
// number of rows whose gate equation fails; public_values[i] = x_i of public row i; p2_row[r] for Qcp rows is v[a[r]]
uint64_t CheckGates(const System& s, const std::vector<Fr>& v, int64_t* first_bad);

}  // namespace scs
}  // namespace gpw
