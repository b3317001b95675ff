#include "scs.h"

#include <stdexcept>
#include <unordered_map>

namespace gpw {
namespace scs {
namespace {

struct FrHash {
  size_t operator()(const Fr& a) const {
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < 8; i++) h = (h ^ a.l[i]) * 1099511628211ull;
    return (size_t)h;
  }
};

struct Builder {
  const fe::API& api;
  System& s;
  std::unordered_map<Fr, uint32_t, FrHash> coeff_ids;
  std::vector<uint32_t> api_coeff;  // api coefficient id -> system coefficient id
  std::vector<std::vector<Chain>> level_chains;
  // result of lowering a linear expression: coeff * var (coeff id C_ZERO: the expression is zero)
  struct Side {
    uint32_t coeff, var;
  };
  std::vector<Side> le_side;
  std::vector<uint8_t> le_done;

  Builder(const fe::API& a, System& sys) : api(a), s(sys) {}

  uint32_t coeff(const Fr& c) {
    auto it = coeff_ids.find(c);
    if (it != coeff_ids.end()) return it->second;
    const uint32_t id = (uint32_t)s.coeffs.size();
    s.coeffs.push_back(c);
    coeff_ids[c] = id;
    return id;
  }

  void gate(uint32_t a, uint32_t b, uint32_t c, uint32_t ql, uint32_t qr, uint32_t qm, uint32_t qo, uint32_t qc, uint8_t qcp = 0) {
    s.a.push_back(a);
    s.b.push_back(b);
    s.c.push_back(c);
    s.ql.push_back(ql);
    s.qr.push_back(qr);
    s.qm.push_back(qm);
    s.qo.push_back(qo);
    s.qc.push_back(qc);
    s.qcp.push_back(qcp);
  }

  // one chain konst + sum of (coefficient id, variable) terms at `level`; returns the variable holding the sum
  uint32_t chain(const uint32_t* cs, const uint32_t* ws, uint32_t n, uint32_t level, uint32_t konst) {
    if (level_chains.size() <= level) level_chains.resize(level + 1);
    Chain ch;
    ch.term_off = (uint32_t)s.chain_wire.size();
    ch.n_terms = n;
    ch.out = s.n_vars;
    ch.konst = konst;
    for (uint32_t j = 0; j < n; j++) {
      s.chain_wire.push_back(ws[j]);
      s.chain_coeff.push_back(cs[j]);
    }
    const uint32_t n_out = n > 1 ? n - 1 : 1;
    if ((uint64_t)s.n_vars + n_out >= 0xfffffff0ull) throw std::runtime_error("scs: too many variables");
    s.n_vars += n_out;
    level_chains[level].push_back(ch);
    if (n == 1) {
      gate(ws[0], 0, ch.out, cs[0], System::C_ZERO, System::C_ZERO, System::C_NEG_ONE, konst);
      return ch.out;
    }
    gate(ws[0], ws[1], ch.out, cs[0], cs[1], System::C_ZERO, System::C_NEG_ONE, konst);
    for (uint32_t j = 2; j < n; j++)
      gate(ch.out + j - 2, ws[j], ch.out + j - 1, System::C_ONE, cs[j], System::C_ZERO, System::C_NEG_ONE, System::C_ZERO);
    return ch.out + n - 2;
  }

  // konst + sum of n terms as one variable: chunks of SCS_CHUNK at `level`, their sums chained one level up
  uint32_t sum(std::vector<uint32_t>& cs, std::vector<uint32_t>& ws, uint32_t level, uint32_t konst) {
    const uint32_t n = (uint32_t)ws.size();
    if (n <= SCS_CHUNK) return chain(cs.data(), ws.data(), n, level, konst);
    std::vector<uint32_t> ncs, nws;
    for (uint32_t lo = 0; lo < n; lo += SCS_CHUNK) {
      const uint32_t m = std::min(SCS_CHUNK, n - lo);
      if (m == 1) {  // a lone last term joins the upper chain as it is
        ncs.push_back(cs[lo]);
        nws.push_back(ws[lo]);
      } else {
        nws.push_back(chain(cs.data() + lo, ws.data() + lo, m, level, lo == 0 ? konst : System::C_ZERO));
        ncs.push_back(System::C_ONE);
      }
    }
    return sum(ncs, nws, level + 1, System::C_ZERO);
  }

  Side lower(uint32_t le) {
    if (le_done[le]) return le_side[le];
    const auto& off = api.LeOffsets();
    const auto& lw = api.LeWires();
    const auto& lc = api.LeCoeffIds();
    const uint32_t k = off[le + 1] - off[le];
    Side r{System::C_ZERO, 0};
    if (k == 1) {
      r = {api_coeff[lc[off[le]]], lw[off[le]]};
    } else if (k >= 2) {
      // terms are sorted by wire id: a constant term (wire 0 = ONE) comes first and rides on qC
      uint32_t t0 = 0, konst = System::C_ZERO;
      if (lw[off[le]] == 0) {
        konst = api_coeff[lc[off[le]]];
        t0 = 1;
      }
      std::vector<uint32_t> cs(k - t0), ws(k - t0);
      for (uint32_t t = t0; t < k; t++) {
        cs[t - t0] = api_coeff[lc[off[le] + t]];
        ws[t - t0] = lw[off[le] + t];
      }
      r = {System::C_ONE, sum(cs, ws, 0, konst)};
    }
    le_side[le] = r;
    le_done[le] = 1;
    return r;
  }
};

}  // namespace

System Build(const fe::API& api) {
  System s;
  Builder b(api, s);
  s.coeffs.clear();
  b.coeff(Fr::zero());
  b.coeff(Fr::one());
  b.coeff(neg(Fr::one()));
  b.api_coeff.resize(api.Coeffs().size());
  for (size_t i = 0; i < api.Coeffs().size(); i++) b.api_coeff[i] = b.coeff(api.Coeffs()[i]);
  s.n_orig = s.n_vars = api.NumWires();
  s.has_commit = api.NumLimbWires() != 0;
  s.commit_wire = api.CommitWire();
  s.committed_lo = api.LimbWireStart();
  s.n_committed = s.has_commit ? api.NumLimbWires() + 65536u : 0u;
  // public rows: a - x_i = 0
  s.public_var.push_back(0);
  for (uint32_t i = 0; i < api.NumPublic(); i++) s.public_var.push_back(1 + i);
  if (s.has_commit) s.public_var.push_back(s.commit_wire);
  for (uint32_t v : s.public_var) b.gate(v, 0, 0, System::C_ONE, System::C_ZERO, System::C_ZERO, System::C_ZERO, System::C_ZERO);
  s.n_public_rows = (uint32_t)s.public_var.size();
  // committed wires: -a + P2 = 0
  for (uint32_t i = 0; i < s.n_committed; i++)
    b.gate(s.committed_lo + i, 0, 0, System::C_NEG_ONE, System::C_ZERO, System::C_ZERO, System::C_ZERO, System::C_ZERO, 1);
  s.n_qcp_rows = s.n_committed;
  // R1CS rows
  const size_t n_le = api.LeOffsets().size() - 1;
  b.le_side.assign(n_le, Builder::Side{System::C_ZERO, 0});
  b.le_done.assign(n_le, 0);
  const auto& cons = api.Constraints();
  for (size_t j = 0; j < cons.size() / 3; j++) {
    const Builder::Side L = b.lower(cons[3 * j]), R = b.lower(cons[3 * j + 1]), O = b.lower(cons[3 * j + 2]);
    const bool lz = L.coeff == System::C_ZERO, rz = R.coeff == System::C_ZERO, oz = O.coeff == System::C_ZERO;
    if (lz || rz) {
      if (oz) continue;  // 0 = 0
      b.gate(0, 0, O.var, System::C_ZERO, System::C_ZERO, System::C_ZERO, O.coeff, System::C_ZERO);  // O = 0
      continue;
    }
    const uint32_t qm = b.coeff(mul(s.coeffs[L.coeff], s.coeffs[R.coeff]));
    const uint32_t qo = oz ? System::C_ZERO : b.coeff(neg(s.coeffs[O.coeff]));
    b.gate(L.var, R.var, oz ? 0u : O.var, System::C_ZERO, System::C_ZERO, qm, qo, System::C_ZERO);
  }
  for (const auto& lvl : b.level_chains) {
    s.level_off.push_back((uint32_t)s.chains.size());
    s.chains.insert(s.chains.end(), lvl.begin(), lvl.end());
  }
  s.level_off.push_back((uint32_t)s.chains.size());
  s.n_gates = (uint32_t)s.a.size();
  s.logN = 2;
  while ((1ull << s.logN) < s.n_gates) s.logN++;
  if (s.logN > 27) throw std::runtime_error("scs: the circuit needs more than 2^27 rows");
  return s;
}

void BuildPermutation(const System& s, std::vector<uint32_t>* sigma) {
  const size_t N = (size_t)1 << s.logN, S = 3 * N;
  auto var_of = [&](size_t slot) -> uint32_t {
    const size_t col = slot / N, row = slot % N;
    if (row >= s.n_gates) return 0;
    return col == 0 ? s.a[row] : col == 1 ? s.b[row] : s.c[row];
  };
  std::vector<uint32_t> start(s.n_vars + 1, 0);
  for (size_t t = 0; t < S; t++) start[var_of(t) + 1]++;
  for (size_t v = 0; v < s.n_vars; v++) start[v + 1] += start[v];
  std::vector<uint32_t> order(S), fill(start.begin(), start.end() - 1);
  for (size_t t = 0; t < S; t++) order[fill[var_of(t)]++] = (uint32_t)t;
  sigma->assign(S, 0);
  for (size_t v = 0; v < s.n_vars; v++) {
    const uint32_t lo = start[v], hi = start[v + 1];
    for (uint32_t i = lo; i < hi; i++) (*sigma)[order[i]] = order[i + 1 < hi ? i + 1 : lo];
  }
}

void ExtendWitness(const System& s, std::vector<Fr>* vp) {
  std::vector<Fr>& v = *vp;
  v.resize(s.n_vars, Fr::zero());
  for (const Chain& ch : s.chains) {
    Fr acc = add(s.coeffs[ch.konst], mul(s.coeffs[s.chain_coeff[ch.term_off]], v[s.chain_wire[ch.term_off]]));
    if (ch.n_terms == 1) v[ch.out] = acc;
    for (uint32_t j = 1; j < ch.n_terms; j++) {
      acc = add(acc, mul(s.coeffs[s.chain_coeff[ch.term_off + j]], v[s.chain_wire[ch.term_off + j]]));
      v[ch.out + j - 1] = acc;
    }
  }
}

uint64_t CheckGates(const System& s, const std::vector<Fr>& v, int64_t* first_bad) {
  uint64_t bad = 0;
  if (first_bad) *first_bad = -1;
  for (uint32_t r = 0; r < s.n_gates; r++) {
    const Fr &va = v[s.a[r]], &vb = v[s.b[r]], &vc = v[s.c[r]];
    Fr t = mul(s.coeffs[s.ql[r]], va);
    t = add(t, mul(s.coeffs[s.qr[r]], vb));
    t = add(t, mul(s.coeffs[s.qm[r]], mul(va, vb)));
    t = add(t, mul(s.coeffs[s.qo[r]], vc));
    t = add(t, s.coeffs[s.qc[r]]);
    if (r < s.n_public_rows) t = sub(t, v[s.public_var[r]]);  // PI(row) = -x_i
    if (s.qcp[r]) t = add(t, v[s.a[r]]);                      // P2(row) = the committed wire
    if (!t.is_zero()) {
      if (!bad && first_bad) *first_bad = r;
      bad++;
    }
  }
  return bad;
}

}  // namespace scs
}  // namespace gpw
