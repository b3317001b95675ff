#include "frontend.h"

#include <algorithm>
#include <istream>
#include <ostream>
#include <type_traits>

namespace gpw {
namespace fe {

Fr fr_from_u64(uint64_t v) {
  Fr a = Fr::zero();
  a.l[0] = (uint32_t)v;
  a.l[1] = (uint32_t)(v >> 32);
  return to_mont(a);
}

Fr fr_from_limbs(const uint64_t l[4]) {
  Fr a;
  memcpy(a.l, l, 32);
  return to_mont(a);
}

void fr_to_limbs(const Fr& a, uint64_t l[4]) {
  Fr c = from_mont(a);
  memcpy(l, c.l, 32);
}

Fr fr_from_dec(const std::string& s) {
  Fr acc = Fr::zero();
  const Fr ten = fr_from_u64(10);
  for (char ch : s) {
    if (ch < '0' || ch > '9') throw std::invalid_argument("fr_from_dec: not a decimal string: " + s);
    acc = add(mul(acc, ten), fr_from_u64((uint64_t)(ch - '0')));
  }
  return acc;
}

API::API() {
  wire_level_.push_back(0);  // ONE
  wire_bool_.push_back(1);
  coeffs_.push_back(Fr::one());       // COEFF_ONE
  coeffs_.push_back(neg(Fr::one()));  // COEFF_NEG_ONE
  coeff_ids_[coeffs_[0]] = 0;
  coeff_ids_[coeffs_[1]] = 1;
  le_off_.push_back(0);
}

uint32_t API::new_wire(uint32_t level) {
  uint32_t w = next_wire_++;
  wire_level_.push_back(level);
  wire_bool_.push_back(0);
  if (level > max_level_) max_level_ = level;
  return w;
}

Variable API::PublicInput() {
  if (inputs_closed_ || n_secret_) throw std::logic_error("public inputs must be declared first");
  n_public_++;
  return wire_var(new_wire(0));
}

Variable API::SecretInput() {
  if (inputs_closed_) throw std::logic_error("inputs already closed");
  n_secret_++;
  return wire_var(new_wire(0));
}

Variable API::wire_var(uint32_t w) const {
  Variable v;
  v.t.push_back({w, Fr::one()});
  return v;
}

Variable API::Const(uint64_t v) const { return ConstFr(fr_from_u64(v)); }

Variable API::ConstFr(const Fr& c) const {
  Variable v;
  if (!c.is_zero()) v.t.push_back({0, c});
  return v;
}

bool API::const_value(const Variable& v, Fr* out) const {
  if (v.t.empty()) {
    *out = Fr::zero();
    return true;
  }
  if (v.t.size() == 1 && v.t[0].wire == 0) {
    *out = v.t[0].coeff;
    return true;
  }
  return false;
}

Variable API::Add(const Variable& a, const Variable& b) const {
  Variable r;
  r.t.reserve(a.t.size() + b.t.size());
  size_t i = 0, j = 0;
  while (i < a.t.size() && j < b.t.size()) {
    if (a.t[i].wire < b.t[j].wire) r.t.push_back(a.t[i++]);
    else if (a.t[i].wire > b.t[j].wire) r.t.push_back(b.t[j++]);
    else {
      Fr c = add(a.t[i].coeff, b.t[j].coeff);
      if (!c.is_zero()) r.t.push_back({a.t[i].wire, c});
      i++, j++;
    }
  }
  while (i < a.t.size()) r.t.push_back(a.t[i++]);
  while (j < b.t.size()) r.t.push_back(b.t[j++]);
  return r;
}

Variable API::Neg(const Variable& a) const {
  Variable r = a;
  for (auto& t : r.t) t.coeff = neg(t.coeff);
  return r;
}

Variable API::Sub(const Variable& a, const Variable& b) const { return Add(a, Neg(b)); }

Variable API::MulConst(const Variable& a, const Fr& c) const {
  Variable r;
  if (c.is_zero()) return r;
  if (c == Fr::one()) return a;
  r.t.reserve(a.t.size());
  for (const auto& t : a.t) r.t.push_back({t.wire, mul(t.coeff, c)});
  return r;
}

uint32_t API::level_of(const Variable& v) const {
  uint32_t l = 0;
  for (const auto& t : v.t) l = std::max(l, wire_level_[t.wire]);
  return l;
}

uint32_t API::intern_coeff(const Fr& c) {
  auto it = coeff_ids_.find(c);
  if (it != coeff_ids_.end()) return it->second;
  uint32_t id = (uint32_t)coeffs_.size();
  coeffs_.push_back(c);
  coeff_ids_[c] = id;
  return id;
}

uint32_t API::intern_le(const Variable& v) {
  if (v.t.size() == 1 && v.t[0].wire == 0 && v.t[0].coeff == Fr::one()) {
    if (le_one_ != NO_LE) return le_one_;
  }
  uint32_t id = (uint32_t)le_off_.size() - 1;
  for (const auto& t : v.t) {
    le_wire_.push_back(t.wire);
    le_coeff_.push_back(intern_coeff(t.coeff));
  }
  le_off_.push_back((uint32_t)le_wire_.size());
  if (v.t.size() == 1 && v.t[0].wire == 0 && v.t[0].coeff == Fr::one()) le_one_ = id;
  return id;
}

void API::add_constraint(const Variable& l, const Variable& r, const Variable& o) {
  cons_.push_back(intern_le(l));
  cons_.push_back(intern_le(r));
  cons_.push_back(intern_le(o));
}

Variable API::Mul(const Variable& a, const Variable& b) {
  Fr c;
  if (const_value(a, &c)) return MulConst(b, c);
  if (const_value(b, &c)) return MulConst(a, c);
  uint32_t lvl = std::max(level_of(a), level_of(b)) + 1;
  uint32_t w = new_wire(lvl);
  uint32_t la = intern_le(a), lb = intern_le(b);
  tape_.push_back({OP_MUL, w, 1, {la, lb, NO_LE}, lvl});
  Variable out = wire_var(w);
  cons_.push_back(la);
  cons_.push_back(lb);
  cons_.push_back(intern_le(out));
  counts_.mul++;
  return out;
}

std::vector<Variable> API::NewHint(Op op, uint32_t nout, const Variable* a, const Variable* b, const Variable* c) {
  uint32_t lvl = 0;
  uint32_t le[3] = {NO_LE, NO_LE, NO_LE};
  const Variable* in[3] = {a, b, c};
  for (int i = 0; i < 3; i++)
    if (in[i]) {
      lvl = std::max(lvl, level_of(*in[i]));
      le[i] = intern_le(*in[i]);
    }
  lvl += 1;
  uint32_t first = next_wire_;
  std::vector<Variable> outs;
  outs.reserve(nout);
  for (uint32_t i = 0; i < nout; i++) outs.push_back(wire_var(new_wire(lvl)));
  tape_.push_back({(uint8_t)op, first, nout, {le[0], le[1], le[2]}, lvl});
  if (op == OP_HINT_MULADD || op == OP_HINT_REDUCE || op == OP_HINT_GLINV || op == OP_HINT_SPLIT)
    for (uint32_t i = 0; i < nout; i++) hint_log_.push_back({(uint8_t)op, first + i});
  switch (op) {
    case OP_HINT_MULADD: counts_.muladd++; break;
    case OP_HINT_REDUCE: counts_.reduce++; break;
    case OP_HINT_GLINV: counts_.glinv++; break;
    case OP_HINT_SPLIT: counts_.split++; break;
    case OP_INVZERO: counts_.invzero++; break;
    case OP_BITS: counts_.bits++; break;
    case OP_DIV: counts_.div++; break;
    case OP_DECOMP: counts_.decomp++; break;
    default: break;
  }
  return outs;
}

Variable API::IsZero(const Variable& a) {
  Fr c;
  if (const_value(a, &c)) return Const(c.is_zero() ? 1 : 0);
  // x = 1/a (0 if a == 0);  m = 1 - a x;  a m = 0      (gnark r1cs builder IsZero)
  Variable x = NewHint(OP_INVZERO, 1, &a)[0];
  Variable na = Neg(a);
  uint32_t lvl = std::max(level_of(a), level_of(x)) + 1;
  uint32_t m = new_wire(lvl);
  wire_bool_[m] = 1;
  Variable one = Const(1);
  uint32_t l0 = intern_le(na), l1 = intern_le(x), l2 = intern_le(one);
  tape_.push_back({OP_MUL, m, 1, {l0, l1, l2}, lvl});  // m = (-a) x + 1
  Variable mv = wire_var(m);
  cons_.push_back(l0);
  cons_.push_back(l1);
  cons_.push_back(intern_le(Sub(mv, one)));
  add_constraint(a, mv, Variable());
  counts_.mul++;
  return mv;
}

void API::AssertIsBoolean(const Variable& b) {
  Fr c;
  if (const_value(b, &c)) {
    if (!(c.is_zero() || c == Fr::one())) throw std::logic_error("AssertIsBoolean: constant is not 0/1");
    return;
  }
  if (b.t.size() == 1 && b.t[0].coeff == Fr::one() && wire_bool_[b.t[0].wire]) return;
  add_constraint(b, Sub(Const(1), b), Variable());
  if (b.t.size() == 1 && b.t[0].coeff == Fr::one()) wire_bool_[b.t[0].wire] = 1;
}

Variable API::Select(const Variable& b, const Variable& i1, const Variable& i2) {
  Fr c;
  if (const_value(b, &c)) {
    if (c == Fr::one()) return i1;
    if (c.is_zero()) return i2;
    throw std::logic_error("Select: constant condition is not boolean");
  }
  AssertIsBoolean(b);
  Variable d = Sub(i1, i2);
  if (d.is_zero()) return i2;
  return Add(Mul(b, d), i2);
}

Variable API::Lookup2(const Variable& b0, const Variable& b1, const Variable& i0, const Variable& i1,
                      const Variable& i2, const Variable& i3) {
  Variable s0 = Select(b0, i1, i0);
  Variable s1 = Select(b0, i3, i2);
  return Select(b1, s1, s0);
}

std::vector<Variable> API::ToBinary(const Variable& v, int n) {
  Fr c;
  if (const_value(v, &c)) {
    uint64_t l[4];
    fr_to_limbs(c, l);
    std::vector<Variable> bits;
    for (int i = 0; i < n; i++) bits.push_back(Const((l[i >> 6] >> (i & 63)) & 1));
    for (int i = n; i < 256; i++)
      if ((l[i >> 6] >> (i & 63)) & 1) throw std::logic_error("ToBinary: constant does not fit");
    return bits;
  }
  std::vector<Variable> bits = NewHint(OP_BITS, (uint32_t)n, &v);
  for (auto& b : bits) AssertIsBoolean(b);
  AssertIsEqual(FromBinary(bits, 0, bits.size()), v);
  return bits;
}

Variable API::FromBinary(const std::vector<Variable>& bits, size_t lo, size_t hi) const {
  Variable acc;
  Fr p = Fr::one();
  for (size_t i = lo; i < hi; i++) {
    acc = Add(acc, MulConst(bits[i], p));
    p = dbl(p);
  }
  return acc;
}

Variable API::DivUnchecked(const Variable& a, const Variable& b) {
  Fr c;
  if (const_value(b, &c)) {
    if (c.is_zero()) throw std::logic_error("DivUnchecked: division by constant zero");
    return MulConst(a, inv(c));
  }
  Variable out = NewHint(OP_DIV, 1, &a, &b)[0];
  add_constraint(out, b, a);
  return out;
}

void API::AssertIsEqual(const Variable& a, const Variable& b) {
  Fr ca, cb;
  if (const_value(a, &ca) && const_value(b, &cb)) {
    if (ca != cb) throw std::logic_error("AssertIsEqual: constants differ");
    return;
  }
  add_constraint(a, Const(1), b);
}

void API::FuseAsMacro(Op op, size_t tape_begin, uint32_t wire_begin, const Variable in[4]) {
  const uint32_t nout = next_wire_ - wire_begin;
  if (nout == 0) return;  // everything folded to constants
  uint32_t expect = wire_begin;
  for (size_t i = tape_begin; i < tape_.size(); i++) {
    if (tape_[i].op != OP_MUL || tape_[i].nout != 1 || tape_[i].out != expect)
      throw std::logic_error("FuseAsMacro: the fused region must consist of multiplications defining consecutive wires");
    expect++;
  }
  if (expect != next_wire_) throw std::logic_error("FuseAsMacro: wires created outside the tape in the fused region");
  counts_.mul -= (tape_.size() - tape_begin);
  tape_.resize(tape_begin);
  uint32_t lvl = 0, le[4];
  for (int i = 0; i < 4; i++) {
    lvl = std::max(lvl, level_of(in[i]));
    le[i] = intern_le(in[i]);
  }
  lvl += 1;
  for (uint32_t w = wire_begin; w < next_wire_; w++) wire_level_[w] = lvl;
  if (lvl > max_level_) max_level_ = lvl;
  Instr m{(uint8_t)op, wire_begin, nout, {le[0], le[1], le[2]}, lvl};
  m.le3 = le[3];
  tape_.push_back(m);
}

uint32_t API::intern_raw_le(const std::vector<Term>& terms) {
  const uint32_t id = (uint32_t)le_off_.size() - 1;
  for (const Term& t : terms) {
    le_wire_.push_back(t.wire);
    le_coeff_.push_back(intern_coeff(t.coeff));
  }
  le_off_.push_back((uint32_t)le_wire_.size());
  return id;
}

void API::BeginFuse() {
  if (fuse_active_) throw std::logic_error("BeginFuse: already capturing");
  fuse_active_ = true;
  fuse_begin_ = tape_.size();
  fuse_marked_.clear();
}

void API::FuseMarkSince(size_t tape_before) {
  if (!fuse_active_) return;
  for (size_t i = tape_before; i < tape_.size(); i++) fuse_marked_.push_back(i);
}

bool API::EndFuse(Op op, const Variable* in, size_t n_in, uint32_t expect_nout) {
  if (!fuse_active_) throw std::logic_error("EndFuse without BeginFuse");
  fuse_active_ = false;
  // inputs: term k of the vector expression = input k (constant -> coefficient on the ONE wire)
  std::vector<Term> vec;
  uint32_t lvl = 0;
  for (size_t k = 0; k < n_in; k++) {
    const Variable& v = in[k];
    if (v.t.empty()) vec.push_back({0, Fr::zero()});
    else if (v.t.size() == 1) vec.push_back(v.t[0]);
    else return false;
    lvl = std::max(lvl, level_of(v));
  }
  std::vector<uint32_t> outs;
  for (size_t i : fuse_marked_) {
    const Instr& m = tape_[i];
    if (i < fuse_begin_ || !(m.op == OP_MUL || m.op == OP_HINT_MULADD || m.op == OP_HINT_REDUCE) || m.outs_off != NO_LE)
      throw std::logic_error("EndFuse: unexpected instruction marked for fusion");
    for (uint32_t k = 0; k < m.nout; k++) outs.push_back(m.out + k);
  }
  if (outs.size() != expect_nout) return false;  // some product folded to a constant: keep the plain tape
  lvl += 1;
  // the macro takes the place of the first instruction of the region (the tape stays in dependency order: the
  // retained instructions - range checks of the hint outputs - read the macro's outputs); the marked ones go
  std::vector<uint8_t> drop(tape_.size() - fuse_begin_, 0);
  for (size_t i : fuse_marked_) drop[i - fuse_begin_] = 1;
  std::vector<Instr> kept;
  for (size_t i = fuse_begin_; i < tape_.size(); i++)
    if (!drop[i - fuse_begin_]) kept.push_back(tape_[i]);
  tape_.resize(fuse_begin_);
  for (uint32_t o : outs) wire_level_[o] = lvl;
  if (lvl > max_level_) max_level_ = lvl;
  Instr m{(uint8_t)op, outs.front(), (uint32_t)outs.size(), {intern_raw_le(vec), NO_LE, NO_LE}, lvl};
  m.outs_off = (uint32_t)macro_outs_.size();
  macro_outs_.insert(macro_outs_.end(), outs.begin(), outs.end());
  tape_.push_back(m);
  // The kept instructions were levelled against the hints' own (per-element) levels; the macro sits at the level of its
  // latest input, which can be later than that. Re-level them in tape (= dependency) order.
  for (Instr& in : kept) {
    uint32_t need = 0;
    const uint32_t les[4] = {in.le[0], in.le[1], in.le[2], in.le3};
    for (uint32_t le : les) {
      if (le == NO_LE) continue;
      for (uint32_t t = le_off_[le]; t < le_off_[le + 1]; t++) need = std::max(need, wire_level_[le_wire_[t]] + 1);
    }
    if (need > in.level) {
      in.level = need;
      for (uint32_t k = 0; k < in.nout; k++) wire_level_[in.out + k] = need;
      if (need > max_level_) max_level_ = need;
    }
    tape_.push_back(in);
  }
  return true;
}

void API::RangeCheckCollect(const Variable& v, int bits) {
  if (bits % 16 != 0) throw std::logic_error("v.bits is not nbBits aligned");  // goldilocks/base.go:433-435
  rc_.push_back({intern_le(v), bits});
}

uint64_t API::NumRangeCheckedLimbs() const {
  uint64_t n = 0;
  for (auto& p : rc_) n += (uint64_t)p.second / 16;
  return n;
}

void API::Finalize() {
  if (finalized_) throw std::logic_error("Finalize called twice");
  finalized_ = true;
  if (rc_.empty()) return;
  // 1. limb decomposition of every collected value; all of it runs as ONE wide level after the main tape
  const uint32_t decomp_level = max_level_ + 1;
  limb_wire_start_ = next_wire_;
  std::vector<Variable> all_limbs;
  all_limbs.reserve(NumRangeCheckedLimbs());
  for (auto& p : rc_) {
    uint32_t k = (uint32_t)p.second / 16;
    uint32_t first = next_wire_;
    Variable acc;
    Fr pw = Fr::one();
    const Fr b16 = fr_from_u64(1u << 16);
    for (uint32_t i = 0; i < k; i++) {
      uint32_t w = new_wire(decomp_level);
      Variable lv = wire_var(w);
      all_limbs.push_back(lv);
      acc.t.push_back({w, pw});
      pw = mul(pw, b16);
    }
    tape_.push_back({OP_DECOMP, first, k, {p.first, NO_LE, NO_LE}, decomp_level});
    counts_.decomp++;
    // recomposition: sum limb_i 2^(16 i) == v
    cons_.push_back(intern_le(acc));
    cons_.push_back(intern_le(Const(1)));
    cons_.push_back(p.first);
  }
  n_limb_wires_ = next_wire_ - limb_wire_start_;
  // 2. multiplicities of the 2^16 table entries
  const uint32_t count_level = decomp_level + 1;
  count_wire_start_ = next_wire_;
  for (uint32_t i = 0; i < 65536; i++) new_wire(count_level);
  tape_.push_back({OP_COUNT, count_wire_start_, 65536, {NO_LE, NO_LE, NO_LE}, count_level});
  // 3. commitment -> challenge X
  commit_level_ = count_level + 1;
  commit_wire_ = new_wire(commit_level_);
  tape_.push_back({OP_COMMIT, commit_wire_, 1, {NO_LE, NO_LE, NO_LE}, commit_level_});
  Variable X = wire_var(commit_wire_);
  // 4. sum_i e_i / (X - i)  ==  sum_j 1 / (X - f_j)
  Variable lhs, rhs;
  Variable one = Const(1);
  for (uint32_t i = 0; i < 65536; i++) {
    Variable e = wire_var(count_wire_start_ + i);
    Variable q = DivUnchecked(e, Sub(X, Const(i)));
    lhs.t.push_back(q.t[0]);
  }
  for (auto& f : all_limbs) {
    Variable q = DivUnchecked(one, Sub(X, f));
    rhs.t.push_back(q.t[0]);
  }
  AssertIsEqual(lhs, rhs);
}

void API::ScheduleALAP() {
  if (tape_.empty()) return;
  const uint32_t L = max_level_;
  std::vector<uint32_t> producer(next_wire_, NO_LE);  // wire -> tape index
  for (uint32_t i = 0; i < tape_.size(); i++)
    for (uint32_t k = 0; k < tape_[i].nout; k++) producer[tape_[i].out_wire(k, macro_outs_)] = i;
  std::vector<uint32_t> alap(tape_.size(), L);
  int64_t count_idx = -1, commit_idx = -1;
  for (uint32_t i = 0; i < tape_.size(); i++) {
    if (tape_[i].op == OP_COUNT) count_idx = i;
    if (tape_[i].op == OP_COMMIT) commit_idx = i;
  }
  for (size_t ii = tape_.size(); ii-- > 0;) {
    const Instr& in = tape_[ii];
    if (in.op == OP_COMMIT && count_idx >= 0) alap[count_idx] = std::min(alap[count_idx], alap[ii] - 1);
    if (in.op == OP_COUNT) continue;  // its inputs (all DECOMP outputs) are handled below
    if (in.op == OP_DECOMP && count_idx >= 0) alap[ii] = std::min(alap[ii], alap[count_idx] - 1);
    for (int j = 0; j < 4; j++) {
      const uint32_t lej = j < 3 ? in.le[j] : in.le3;
      if (lej == NO_LE) continue;
      for (uint32_t k = le_off_[lej]; k < le_off_[lej + 1]; k++) {
        uint32_t p = producer[le_wire_[k]];
        if (p != NO_LE) alap[p] = std::min(alap[p], alap[ii] - 1);
      }
    }
  }
  (void)commit_idx;
  for (uint32_t i = 0; i < tape_.size(); i++) {
    if (alap[i] < tape_[i].level) throw std::logic_error("ALAP level below ASAP level");
    tape_[i].level = alap[i];
    for (uint32_t k = 0; k < tape_[i].nout; k++) wire_level_[tape_[i].out_wire(k, macro_outs_)] = alap[i];
    if (tape_[i].op == OP_COMMIT) commit_level_ = alap[i];
  }
}

void API::ScheduleSpineAndTail() {
  if (scheduled_) return;
  scheduled_ = true;
  if (tape_.empty()) return;
  const size_t n = tape_.size();
  std::vector<uint32_t> producer(next_wire_, NO_LE);
  for (uint32_t i = 0; i < n; i++)
    for (uint32_t k = 0; k < tape_[i].nout; k++) producer[tape_[i].out_wire(k, macro_outs_)] = i;
  // 1. height = longest path (in instructions) from an instruction down to a sink of the tape. The spine of the
  //    verifier (sponge -> FRI) has heights in the tens of thousands; the side branches every hinted operation
  //    sprouts (SplitLimbs -> IsZero inverse -> select, limb decomposition -> histogram -> commitment -> divisions)
  //    end within a handful of steps. Everything of small height is "tail".
  constexpr uint32_t TAIL_HEIGHT = 7;
  std::vector<uint32_t> height(n, 0);
  int64_t cnt_i = -1, com_i = -1;
  for (uint32_t i = 0; i < n; i++) {
    if (tape_[i].op == OP_COUNT) cnt_i = i;
    if (tape_[i].op == OP_COMMIT) com_i = i;
  }
  for (size_t ii = n; ii-- > 0;) {
    const Instr& in = tape_[ii];
    if (in.op == OP_COUNT && com_i >= 0) height[ii] = std::max(height[ii], height[com_i] + 1);   // commitment covers the counts
    if (in.op == OP_DECOMP && cnt_i >= 0) height[ii] = std::max(height[ii], height[cnt_i] + 1);  // histogram reads the limbs
    for (int j = 0; j < 4; j++) {
      const uint32_t lej = j < 3 ? in.le[j] : in.le3;
      if (lej == NO_LE) continue;
      for (uint32_t k = le_off_[lej]; k < le_off_[lej + 1]; k++) {
        uint32_t p = producer[le_wire_[k]];
        if (p != NO_LE) height[p] = std::max(height[p], height[ii] + 1);
      }
    }
  }
  std::vector<uint8_t> tail(n, 0);
  for (size_t i = 0; i < n; i++) tail[i] = height[i] <= TAIL_HEIGHT ? 1 : 0;
  // 2. levels: spine keeps its ASAP level (parallel branches stay aligned, so the 28 FRI queries execute their
  //    expensive inversions in the same level); tail instructions are levelled among themselves after the spine.
  uint32_t spine_max = 0;
  for (size_t i = 0; i < n; i++)
    if (!tail[i]) spine_max = std::max(spine_max, tape_[i].level);
  std::vector<uint32_t> tl(n, 0);
  uint32_t max_level = spine_max;
  int64_t count_idx = -1;
  uint32_t decomp_max = 0;
  for (size_t i = 0; i < n; i++) {
    if (!tail[i]) continue;
    const Instr& in = tape_[i];
    uint32_t lvl = spine_max + 1;
    for (int j = 0; j < 4; j++) {
      const uint32_t lej = j < 3 ? in.le[j] : in.le3;
      if (lej == NO_LE) continue;
      for (uint32_t k = le_off_[lej]; k < le_off_[lej + 1]; k++) {
        uint32_t p = producer[le_wire_[k]];
        if (p != NO_LE && tail[p]) lvl = std::max(lvl, tl[p] + 1);
      }
    }
    if (in.op == OP_DECOMP) decomp_max = std::max(decomp_max, lvl);
    if (in.op == OP_COUNT) {
      lvl = std::max(lvl, decomp_max + 1);
      count_idx = (int64_t)i;
    }
    if (in.op == OP_COMMIT && count_idx >= 0) lvl = std::max(lvl, tl[count_idx] + 1);
    tl[i] = lvl;
    max_level = std::max(max_level, lvl);
  }
  for (size_t i = 0; i < n; i++) {
    if (tail[i]) tape_[i].level = tl[i];
    for (uint32_t k = 0; k < tape_[i].nout; k++) wire_level_[tape_[i].out_wire(k, macro_outs_)] = tape_[i].level;
    if (tape_[i].op == OP_COMMIT) commit_level_ = tape_[i].level;
  }
  max_level_ = max_level;
}

}  // namespace fe
}  // namespace gpw

// ---- compile cache ------------------------------------------------------------------------------------------------------
namespace gpw {
namespace fe {
namespace {
constexpr uint64_t CACHE_MAGIC = 0x3143575047ull;  // "GPWC1"
template <class T>
void put_vec(std::ostream& os, const std::vector<T>& v) {
  static_assert(std::is_trivially_copyable<T>::value, "raw dump");
  const uint64_t n = v.size();
  os.write(reinterpret_cast<const char*>(&n), 8);
  if (n) os.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(n * sizeof(T)));
}
template <class T>
void get_vec(std::istream& is, std::vector<T>& v) {
  uint64_t n = 0;
  is.read(reinterpret_cast<char*>(&n), 8);
  if (!is || n > (1ull << 33) / sizeof(T)) throw std::runtime_error("circuit cache: bad vector length");
  v.resize(n);
  if (n) is.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(n * sizeof(T)));
  if (!is) throw std::runtime_error("circuit cache: truncated");
}
// (std::pair is not trivially copyable for the type system although its layout is plain: go through two flat arrays)
template <class A, class B>
void put_pairs(std::ostream& os, const std::vector<std::pair<A, B>>& v) {
  std::vector<A> a(v.size());
  std::vector<B> b(v.size());
  for (size_t i = 0; i < v.size(); i++) {
    a[i] = v[i].first;
    b[i] = v[i].second;
  }
  put_vec(os, a);
  put_vec(os, b);
}
template <class A, class B>
void get_pairs(std::istream& is, std::vector<std::pair<A, B>>& v) {
  std::vector<A> a;
  std::vector<B> b;
  get_vec(is, a);
  get_vec(is, b);
  if (a.size() != b.size()) throw std::runtime_error("circuit cache: pair arrays differ in length");
  v.resize(a.size());
  for (size_t i = 0; i < a.size(); i++) v[i] = {a[i], b[i]};
}
template <class T>
void put_pod(std::ostream& os, const T& v) {
  os.write(reinterpret_cast<const char*>(&v), sizeof(T));
}
template <class T>
void get_pod(std::istream& is, T& v) {
  is.read(reinterpret_cast<char*>(&v), sizeof(T));
  if (!is) throw std::runtime_error("circuit cache: truncated");
}
}  // namespace

void API::Serialize(std::ostream& os) const {
  put_pod(os, CACHE_MAGIC);
  const uint32_t sizes[2] = {(uint32_t)sizeof(Instr), (uint32_t)sizeof(Fr)};
  put_pod(os, sizes);
  const uint32_t head[12] = {next_wire_, n_public_, n_secret_, max_level_, commit_level_, limb_wire_start_, n_limb_wires_,
                             count_wire_start_, commit_wire_, le_one_, (uint32_t)finalized_, (uint32_t)scheduled_};
  put_pod(os, head);
  put_pod(os, counts_);
  put_vec(os, tape_);
  put_vec(os, cons_);
  put_vec(os, le_off_);
  put_vec(os, le_wire_);
  put_vec(os, le_coeff_);
  put_vec(os, coeffs_);
  put_vec(os, macro_outs_);
  put_pairs(os, hint_log_);
  put_pairs(os, rc_);
}

void API::Deserialize(std::istream& is) {
  uint64_t magic = 0;
  get_pod(is, magic);
  uint32_t sizes[2];
  get_pod(is, sizes);
  if (magic != CACHE_MAGIC || sizes[0] != sizeof(Instr) || sizes[1] != sizeof(Fr)) throw std::runtime_error("circuit cache: not a gpw circuit blob of this build");
  uint32_t head[12];
  get_pod(is, head);
  next_wire_ = head[0];
  n_public_ = head[1];
  n_secret_ = head[2];
  max_level_ = head[3];
  commit_level_ = head[4];
  limb_wire_start_ = head[5];
  n_limb_wires_ = head[6];
  count_wire_start_ = head[7];
  commit_wire_ = head[8];
  le_one_ = head[9];
  finalized_ = head[10] != 0;
  scheduled_ = head[11] != 0;
  inputs_closed_ = true;
  get_pod(is, counts_);
  get_vec(is, tape_);
  get_vec(is, cons_);
  get_vec(is, le_off_);
  get_vec(is, le_wire_);
  get_vec(is, le_coeff_);
  get_vec(is, coeffs_);
  get_vec(is, macro_outs_);
  get_pairs(is, hint_log_);
  get_pairs(is, rc_);
  // consistency: every index stays inside its table
  if (le_off_.empty() || le_off_.back() != le_wire_.size() || le_wire_.size() != le_coeff_.size() || cons_.size() % 3)
    throw std::runtime_error("circuit cache: inconsistent tables");
  for (uint32_t le : cons_)
    if (le + 1 >= le_off_.size()) throw std::runtime_error("circuit cache: constraint refers to a missing linear expression");
  for (uint32_t w : le_wire_)
    if (w >= next_wire_) throw std::runtime_error("circuit cache: wire id out of range");
  for (uint32_t c : le_coeff_)
    if (c >= coeffs_.size()) throw std::runtime_error("circuit cache: coefficient id out of range");
  wire_level_.clear();
  wire_bool_.clear();
  coeff_ids_.clear();
}

}  // namespace fe
}  // namespace gpw
