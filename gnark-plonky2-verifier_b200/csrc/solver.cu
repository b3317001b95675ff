// Witness synthesis on the GPU: compiles the verifier circuit (host, csrc/host/*) into a levelled solver
// tape and replays it on the device. Replaces gnark's frontend.Compile (benchmark.go:55) and the
// constraint solver + the reference's four hint functions inside groth16.Prove (benchmark.go:249,
// goldilocks/base.go:223-359; SURVEY 3.1 "r1cs.Solve").
//
// Execution model. After ALAP scheduling the tape has a long, narrow sequential spine (~37 k levels, median
// ONE instruction per level: the Fiat-Shamir sponge, then FRI) and a few very wide final levels (range-check
// splits, IsZero inverses, limb decompositions, the log-derivative divisions: ~4.2 M of the 5.3 M
// instructions).
//   * narrow levels: ONE persistent CTA per proof walks the levels with a block barrier between them - no
//     kernel launch per level; a batch of proofs runs one CTA (= one SM) per proof, independent of each other;
//   * wide levels: ordinary grid launches, (instruction, proof) parallel across all SMs.
// Wires live in HBM as Fr elements in Montgomery form (gnark's in-memory form), one contiguous vector per
// proof; linear expressions are evaluated on the fly from CSR-like (wire, coefficient-id) term lists.
#include <algorithm>
#include <fstream>
#include <sstream>

#include "common.cuh"
#include "ff.cuh"
#include "tma.cuh"
#include "gl.cuh"
#include "host/frontend.h"
#include "host/gadgets.h"
#include "poseidon_bn254_macro.cuh"
#include "poseidon_gl_macro.cuh"
#include "poseidon_constants.inc"

namespace gpw {

using fe::NO_LE;

struct DInstr {
  uint32_t op_nout;  // op | nout << 8
  uint32_t out;
  uint32_t le[4];
};

struct DevCircuit {
  const DInstr* instr;
  const uint32_t* level_off;
  const uint32_t* le_off;
  const uint32_t* le_wire;
  const uint32_t* le_coeff;
  const Fr* coeffs;
  const uint32_t* cons;
  uint32_t n_wires, n_cons, n_levels;
  uint32_t limb_start, n_limbs, count_start, commit_wire;
  // linear expressions with thousands of terms (the two sides of the log-derivative identity: 65 536 and ~2.5 M
  // terms) are reduced by a whole CTA each instead of by the single thread that owns their R1CS row
  uint32_t n_long;
  uint32_t long_le[8];
  Fr* long_val;  // per-proof scratch is not needed: evaluated right before use on the stream
  const Fr* bn_tables;  // Poseidon-BN254 constants: C[88] | S[392] | M[16] | P[16], Montgomery
  const uint64_t* gl_tables;   // Poseidon-Goldilocks constants, layout of glm::T_* (poseidon_gl_macro.cuh)
  const uint32_t* macro_outs;  // output wire lists of the OP_POSEIDON_GL instructions (DInstr.out = offset into it)
};
constexpr uint32_t LONG_LE_TERMS = 2048;

enum SegKind { SEG_NARROW, SEG_WIDE, SEG_COUNT, SEG_COMMIT, SEG_POSEIDON };
struct Segment {
  SegKind kind;
  uint32_t lo, hi;  // level range [lo, hi)
  uint32_t stream_off = 0, first_words = 0;  // SEG_NARROW: location of its staged chunk stream
  // SEG_WIDE: the level's instruction range cut into runs of (batched Fr inversions | Poseidon-BN254 macros | everything else)
  struct Part {
    uint32_t s, t;
    bool inv;
    bool poseidon = false;
  };
  std::vector<Part> parts;
  // SEG_POSEIDON: the Poseidon-BN254 macro instructions [ps, pt) of level `lo`, run by the four-lane grid kernel BEFORE the
  // narrow segment that starts at the same level; that narrow segment leaves them out of its first level (skip_s, skip_t).
  uint32_t ps = 0, pt = 0;
  uint32_t skip_s = 0, skip_t = 0;
};
// a level's Poseidon-BN254 macros leave the spine CTA for the grid kernel from this many permutations on (the spine CTA
// runs one permutation per thread: 784 dependent multiplications at one warp's pace, whatever their number)
constexpr uint32_t POSEIDON4_MIN = 16;

// levels with at least this many instructions run as grid launches across all SMs, narrower ones on the proof's
// persistent spine CTA (GPW_WIDE_THRESHOLD overrides it at circuit-compile time, for experiments). Measured on
// testdata/step: 2048 -> solve phase 1 45.8 ms instead of 47.7 ms, 512 -> 69.7 ms: below a few thousand instructions the
// staged spine (tape and operands in shared memory, no launch per level) beats a grid launch with cold operands.
constexpr uint32_t WIDE_THRESHOLD = 8192;
static uint32_t wide_threshold() {
  const char* e = getenv("GPW_WIDE_THRESHOLD");
  const long v = e ? atol(e) : 0;
  return v >= 64 ? (uint32_t)v : WIDE_THRESHOLD;
}
constexpr int NARROW_THREADS = 256;  // upper bound (launch bounds); the spine launches spine_threads() of them

__device__ __forceinline__ Fr ld_w(const Fr* p) {
  Fr r;
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint4* d = reinterpret_cast<uint4*>(&r);
  d[0] = s[0];
  d[1] = s[1];
  return r;
}
__device__ __forceinline__ void st_w(Fr* p, const Fr& v) {
  const uint4* s = reinterpret_cast<const uint4*>(&v);
  uint4* d = reinterpret_cast<uint4*>(p);
  d[0] = s[0];
  d[1] = s[1];
}

// A linear expression as (wire, coefficient-id) term list. The terms live either in the global CSR arrays
// (stride 1, separate arrays) or inline in a staged tape chunk in shared memory (stride 2, interleaved).
// In the staged stream a term's location is either a wire id (value in HBM) or, with RING_FLAG set, a slot of the
// CTA's shared-memory ring of recently produced wire values (assigned at compile time, see finish_compile).
constexpr uint32_t RING_FLAG = 0x80000000u;
constexpr uint32_t RING_SLOTS = 2048;  // 64 KB of shared memory

struct LeRef {
  const uint32_t* wires;
  const uint32_t* cids;
  uint32_t n, stride;
  bool present;
  const Fr* ring;  // nullptr on the CSR path
};

__device__ __forceinline__ Fr ld_term(const Fr* W, const Fr* ring, uint32_t loc) {
  return (loc & RING_FLAG) ? ld_w(ring + (loc & (RING_SLOTS - 1))) : ld_w(W + loc);
}

// sum coeff_k * w_k. The operand loads of a batch are issued together (independent LDGs in flight) before any
// arithmetic: on the sequential spine of the tape the latency of these loads is the critical path.
// (A real function call, not inlined: exec_op has ~25 call sites and the body holds the unrolled multiplications.)
__device__ __noinline__ Fr eval_ref(const DevCircuit& c, const Fr* W, const LeRef& r) {
  Fr acc = Fr::zero();
  constexpr int BATCH = 4;
  for (uint32_t k0 = 0; k0 < r.n; k0 += BATCH) {
    Fr v[BATCH];
    uint32_t cid[BATCH];
    bool is_const[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; j++)
      if (k0 + j < r.n) {
        cid[j] = r.cids[(k0 + j) * r.stride];
        const uint32_t loc = r.wires[(k0 + j) * r.stride];
        is_const[j] = loc == 0;  // wire 0 is the constant one: the term's value is the coefficient itself
        v[j] = is_const[j] ? ld_w(c.coeffs + cid[j]) : ld_term(W, r.ring, loc);
      }
#pragma unroll
    for (int j = 0; j < BATCH; j++)
      if (k0 + j < r.n) {
        if (cid[j] == 0 || is_const[j]) acc = add(acc, v[j]);
        else if (cid[j] == 1) acc = sub(acc, v[j]);
        else acc = add(acc, mul(ld_w(c.coeffs + cid[j]), v[j]));
      }
  }
  return acc;
}

__device__ __forceinline__ LeRef csr_ref(const DevCircuit& c, uint32_t le) {
  if (le == NO_LE) return {nullptr, nullptr, 0, 1, false, nullptr};
  const uint32_t s = c.le_off[le];
  return {c.le_wire + s, c.le_coeff + s, c.le_off[le + 1] - s, 1, true, nullptr};
}

// where an instruction's outputs go: always HBM, plus the ring on the staged path
struct OutRef {
  Fr* W;
  uint32_t out;
  Fr* ring;
  uint32_t slot;
  const uint32_t* outs = nullptr;  // scattered outputs (OP_POSEIDON_GL): wire id of output k
  __device__ __forceinline__ void put(uint32_t k, const Fr& v) const {
    st_w(W + (outs ? outs[k] : out + k), v);
    if (ring) st_w(ring + ((slot + k) & (RING_SLOTS - 1)), v);
  }
};

// value of term k of a "vector" expression (macro inputs): coefficient * wire, or the coefficient itself on the ONE wire
__device__ __forceinline__ Fr eval_term(const DevCircuit& c, const Fr* W, const LeRef& r, uint32_t k) {
  const uint32_t cid = r.cids[k * r.stride];
  const uint32_t loc = r.wires[k * r.stride];
  if (loc == 0) return ld_w(c.coeffs + cid);
  const Fr v = ld_term(W, r.ring, loc);
  if (cid == 0) return v;
  if (cid == 1) return neg(v);
  return mul(ld_w(c.coeffs + cid), v);
}


__device__ __forceinline__ Fr eval_le(const DevCircuit& c, const Fr* W, uint32_t le) { return eval_ref(c, W, csr_ref(c, le)); }

__device__ __forceinline__ void to_u64x4(const Fr& mont, uint64_t x[4]) {
  Fr a = from_mont(mont);
#pragma unroll
  for (int i = 0; i < 4; i++) x[i] = (uint64_t)a.l[2 * i] | ((uint64_t)a.l[2 * i + 1] << 32);
}
__device__ __forceinline__ Fr from_u64x4(const uint64_t x[4]) {
  Fr a;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a.l[2 * i] = (uint32_t)x[i];
    a.l[2 * i + 1] = (uint32_t)(x[i] >> 32);
  }
  return to_mont(a);
}
__device__ __forceinline__ Fr from_u64(uint64_t v) {
  uint64_t x[4] = {v, 0, 0, 0};
  return from_u64x4(x);
}
__device__ __forceinline__ Fr fr_from_u192(const glm::U192& v) {
  const uint64_t x[4] = {v.l[0], v.l[1], v.l[2], 0};
  return from_u64x4(x);
}

// error codes written to err[proof] (first error wins)
constexpr int ERR_MULADD = 1, ERR_GLINV = 2, ERR_SPLIT = 3, ERR_BITS = 4, ERR_DECOMP = 5, ERR_DIV0 = 6;

__device__ __forceinline__ bool ref_is_const(const LeRef& r) { return r.n == 0 || (r.n == 1 && r.wires[0] == 0); }

__device__ void exec_op(const DevCircuit& c, const Fr* W, uint32_t op_nout, const OutRef& O, const LeRef& A, const LeRef& B,
                        const LeRef& Cc, const LeRef& D, int* err, uint32_t* hist) {
  const uint32_t op = op_nout & 0xffu, nout = op_nout >> 8;
  switch (op) {
    case fe::OP_MUL: {
      const Fr a = eval_ref(c, W, A);
      const bool same = B.wires == A.wires && B.n == A.n;  // x * x (S-boxes): evaluate the operand once
      Fr r = same ? sqr_chain(a) : mul(a, eval_ref(c, W, B));
      if (Cc.present) r = add(r, eval_ref(c, W, Cc));
      O.put(0,r);
      break;
    }
    case fe::OP_HINT_MULADD: {
      uint64_t a[4], b[4], d[4];
      to_u64x4(eval_ref(c, W, A), a);
      to_u64x4(eval_ref(c, W, B), b);
      to_u64x4(eval_ref(c, W, Cc), d);
      if ((a[1] | a[2] | a[3] | b[1] | b[2] | b[3] | d[1] | d[2] | d[3]) || a[0] >= gl::P || b[0] >= gl::P || d[0] >= gl::P) {
        atomicCAS(err, 0, ERR_MULADD);  // goldilocks/base.go:228-232 panics
        return;
      }
      uint64_t q, r;
      gl::mul_add_hint(a[0], b[0], d[0], q, r);
      O.put(0,from_u64(q));
      O.put(1,from_u64(r));
      break;
    }
    case fe::OP_HINT_REDUCE: {
      uint64_t x[4], q[4], r;
      to_u64x4(eval_ref(c, W, A), x);
      gl::reduce_hint(x, q, r);
      O.put(0,from_u64x4(q));
      O.put(1,from_u64(r));
      break;
    }
    case fe::OP_HINT_GLINV: {
      uint64_t x[4];
      to_u64x4(eval_ref(c, W, A), x);
      if ((x[1] | x[2] | x[3]) || x[0] >= gl::P) {
        atomicCAS(err, 0, ERR_GLINV);
        return;
      }
      O.put(0,from_u64(gl::inverse(x[0])));
      break;
    }
    case fe::OP_HINT_SPLIT: {
      uint64_t x[4];
      to_u64x4(eval_ref(c, W, A), x);
      if ((x[1] | x[2] | x[3]) || x[0] >= gl::P) {
        atomicCAS(err, 0, ERR_SPLIT);  // goldilocks/base.go:347-349 returns an error
        return;
      }
      O.put(0,from_u64(x[0] >> 32));
      O.put(1,from_u64(x[0] & 0xffffffffull));
      break;
    }
    case fe::OP_INVZERO: O.put(0,inv_euclid(eval_ref(c, W, A))); break;  // (shift-and-subtract: ~3x the pace of the Fermat chain on a lone warp)
    case fe::OP_BITS: {
      uint64_t x[4];
      to_u64x4(eval_ref(c, W, A), x);
      const Fr one = Fr::one(), zero = Fr::zero();
      for (uint32_t i = 0; i < nout; i++) O.put(i,((x[i >> 6] >> (i & 63)) & 1ull) ? one : zero);
      for (uint32_t i = nout; i < 256; i++)
        if ((x[i >> 6] >> (i & 63)) & 1ull) atomicCAS(err, 0, ERR_BITS);
      break;
    }
    case fe::OP_DIV: {
      Fr d = eval_ref(c, W, B);
      if (d.is_zero()) atomicCAS(err, 0, ERR_DIV0);
      O.put(0,mul(eval_ref(c, W, A), inv_euclid(d)));
      break;
    }
    case fe::OP_DECOMP: {
      uint64_t x[4];
      to_u64x4(eval_ref(c, W, A), x);
      for (uint32_t i = 0; i < nout; i++) {
        uint32_t v = (uint32_t)((x[(16 * i) >> 6] >> ((16 * i) & 63)) & 0xffffull);
        O.put(i,from_u64(v));
        atomicAdd(&hist[v], 1u);
      }
      for (uint32_t i = 16 * nout; i < 256; i += 16)
        if ((x[i >> 6] >> (i & 63)) & 0xffffull) atomicCAS(err, 0, ERR_DECOMP);  // value exceeds its range: unsatisfiable
      break;
    }
    case fe::OP_POSEIDON_BN254: {
      Fr st[4] = {eval_ref(c, W, A), eval_ref(c, W, B), eval_ref(c, W, Cc), eval_ref(c, W, D)};
      const bool isc[4] = {ref_is_const(A), ref_is_const(B), ref_is_const(Cc), ref_is_const(D)};
      Bn254PoseidonTables T{c.bn_tables, c.bn_tables + 88, c.bn_tables + 88 + 392, c.bn_tables + 88 + 392 + 16};
      uint32_t idx = 0;
      poseidon_bn254_trace(st, isc, T, [&](const Fr& v) { O.put(idx++, v); });
      break;
    }
    case fe::OP_POSEIDON_GL: {
      // one thread, sequential (wide levels / the unstaged debugging walker); the staged spine runs the warp form
      uint64_t st[12];
      bool bad = false;
      for (uint32_t k = 0; k < 12; k++) {
        uint64_t x[4];
        to_u64x4(eval_term(c, W, A, k), x);
        bad |= (x[1] | x[2] | x[3]) != 0 || x[0] >= gl::P;
        st[k] = x[0];
      }
      if (bad) {
        atomicCAS(err, 0, ERR_MULADD);  // the permutation starts with gl.Add = MulAddHint: goldilocks/base.go:228-232
        return;
      }
      glm::trace_seq(st, c.gl_tables, [&](uint32_t slot, const glm::U192& v) { O.put(slot, fr_from_u192(v)); });
      break;
    }
    default: break;
  }
}

__device__ __forceinline__ void exec_instr(const DevCircuit& c, Fr* W, const DInstr& in, int* err, uint32_t* hist) {
  OutRef O{W, in.out, nullptr, 0};
  if ((in.op_nout & 0xffu) == fe::OP_POSEIDON_GL) O.outs = c.macro_outs + in.out;
  exec_op(c, W, in.op_nout, O, csr_ref(c, in.le[0]), csr_ref(c, in.le[1]), csr_ref(c, in.le[2]), csr_ref(c, in.le[3]), err, hist);
}

// ---- staged narrow tape ---------------------------------------------------------------------------------------------
// The narrow levels are re-encoded as a flat stream of CHUNKS (u32 words), each self-contained:
//   [0] n_instr   [1] words of the NEXT chunk (0 = last)   [2] 1 if a level ends with this chunk   [3] flags
//   [4 .. 4+n_instr) word offset of each instruction record inside the chunk
//   records: op|nout<<8, out, nA, nB, nC, ring slot, nD, then (wire, coeff-id) pairs of A, B, C, D
//   (a Poseidon-Goldilocks macro sits alone in its chunk, flag CHUNK_FLAG_GL_MACRO, its 1992 output wire ids after A)
// so one contiguous copy brings everything an instruction needs except the wire values themselves. The CTA keeps
// two chunk buffers in shared memory and prefetches chunk i+1 with ONE TMA bulk copy (cp.async.bulk + mbarrier) while it
// executes chunk i: the only exposed global-memory latency per level is the load of the operand wires.
constexpr uint32_t CHUNK_MAX_WORDS = 12288;  // 48 KB per buffer
// + the Poseidon-Goldilocks macro's trace (1992 integers of 192 bits) and its 2 x 12-word exchange buffers
constexpr size_t GLM_TRACE_WORDS64 = (size_t)glm::N_OUT * 3 + 24;
constexpr size_t STAGED_SMEM_BYTES = 2 * CHUNK_MAX_WORDS * 4 + RING_SLOTS * sizeof(Fr) + GLM_TRACE_WORDS64 * 8 + 16;  // 96 + 64 + 47 KB + 2 mbarriers
constexpr uint32_t CHUNK_FLAG_GL_MACRO = 1;  // chunk header word [3]: the chunk is ONE OP_POSEIDON_GL instruction
// The spine CTA is latency bound and shares nothing: when other proofs' MSM / NTT kernels run next to it (several
// proofs in flight, wrap.cu) their warps would saturate the SM's IMAD pipe and stretch every level of the spine. It
// therefore asks for the SM's whole shared memory, which keeps every kernel that uses shared memory off its SM.
constexpr size_t SPINE_EXCLUSIVE_SMEM_BYTES = 227 * 1024;
static int spine_threads() {
  static const int v = [] {
    const char* e = getenv("GPW_SPINE_THREADS");
    int t = e ? atoi(e) : NARROW_THREADS;
    return (t >= 32 && t <= NARROW_THREADS && t % 32 == 0) ? t : NARROW_THREADS;
  }();
  return v;
}
static size_t spine_smem_bytes() {
  static const size_t v = getenv("GPW_SPINE_SHARED_SM") ? STAGED_SMEM_BYTES : SPINE_EXCLUSIVE_SMEM_BYTES;
  return v;
}

// ---- TMA (1-D bulk) + mbarrier: one elected thread starts the copy of a whole chunk (up to 24 KB, contiguous in the
// stream), the copy engine moves it while all 256 threads work on the current chunk, and completion is signalled on a
// shared-memory mbarrier by byte count - no thread spends instructions on the transfer (the 16-byte cp.async loop this
// replaces cost every thread ~6 copies + a commit per chunk, ~1 000 chunks per proof on a latency-bound CTA).
// (helpers: tma.cuh)

__global__ void __launch_bounds__(NARROW_THREADS)
    k_tape_staged(DevCircuit c, const uint32_t* __restrict__ stream, uint32_t first_words, Fr* __restrict__ wires, size_t wire_stride,
                  int* __restrict__ err, uint32_t* __restrict__ hist, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) uint32_t dyn_smem[];
  uint32_t(*buf)[CHUNK_MAX_WORDS] = reinterpret_cast<uint32_t(*)[CHUNK_MAX_WORDS]>(dyn_smem);
  Fr* ring = reinterpret_cast<Fr*>(dyn_smem + 2 * CHUNK_MAX_WORDS);
  uint64_t* glm_io = reinterpret_cast<uint64_t*>(dyn_smem + 2 * CHUNK_MAX_WORDS + RING_SLOTS * 8);  // [0,12) inputs, [12,24) exchange
  uint64_t* glm_trace = glm_io + 24;
  Fr* W = wires + (size_t)blockIdx.x * wire_stride;
  int* e = err + blockIdx.x;
  uint32_t* h = hist + (size_t)blockIdx.x * 65536;
  const uint32_t* src = stream;
  uint32_t words = first_words;
  uint64_t* chunk_bar = glm_trace + (size_t)glm::N_OUT * 3;  // one mbarrier per chunk buffer (dynamic shared memory: the kernel
                                                             // asks for the SM's whole 227 KB, static storage would not fit on top)
  uint32_t bar_parity[2] = {0, 0};
  if (threadIdx.x == 0) {
    mbar_init(&chunk_bar[0], 1);
    mbar_init(&chunk_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&chunk_bar[0], words * 4);
    tma_bulk_load(&buf[0][0], src, words * 4, &chunk_bar[0]);
  }
  int cur = 0;
  // development profile (GPW_SPINE_PROFILE): cycles of thread 0 spent [0] waiting for the chunk, [1] in macro input fetch,
  // [2] in the native permutation, [3] in the macro's conversion + stores, [4] in ordinary chunks; [5] macros, [6] chunks
  unsigned long long pc[7] = {0, 0, 0, 0, 0, 0, 0};
  long long t_prev = clock64();
  auto lap = [&](int k) {
    const long long now = clock64();
    pc[k] += (unsigned long long)(now - t_prev);
    t_prev = now;
  };
  while (words) {
    mbar_wait(&chunk_bar[cur], bar_parity[cur]);  // chunk `cur` has landed (the TMA's byte count completed the phase)
    bar_parity[cur] ^= 1u;
    __syncthreads();  // wires written by the previous chunk are visible; everyone has left the buffer about to be refilled
    lap(0);
    const uint32_t* ch = buf[cur];
    const uint32_t n_instr = ch[0], next_words = ch[1];
    const uint32_t* next_src = src + words;
    if (threadIdx.x == 0 && next_words) {  // prefetch chunk i+1 into the other buffer while chunk i executes
      mbar_expect_tx(&chunk_bar[cur ^ 1], next_words * 4);
      tma_bulk_load(&buf[cur ^ 1][0], next_src, next_words * 4, &chunk_bar[cur ^ 1]);
    }
    // one instruction record of the chunk
    auto run_record = [&](uint32_t i) {
      const uint32_t* rec = ch + ch[4 + i];
      const uint32_t nA = rec[2], nB = rec[3], nC = rec[4], nD = rec[6];
      const uint32_t* t = rec + 7;
      LeRef A{t, t + 1, nA & 0x7fffffffu, 2, (nA >> 31) != 0, ring};
      t += 2 * (nA & 0x7fffffffu);
      LeRef B{t, t + 1, nB & 0x3fffffffu, 2, (nB >> 31) != 0, ring};
      if (nB & 0x40000000u) B = A;  // encoder: B is the same expression as A, its terms are not repeated
      else t += 2 * (nB & 0x3fffffffu);
      LeRef Cc{t, t + 1, nC & 0x7fffffffu, 2, (nC >> 31) != 0, ring};
      t += 2 * (nC & 0x7fffffffu);
      LeRef D{t, t + 1, nD & 0x7fffffffu, 2, (nD >> 31) != 0, ring};
      exec_op(c, W, rec[0], OutRef{W, rec[1], ring, rec[5]}, A, B, Cc, D, e, h);
    };
    if (ch[3] == CHUNK_FLAG_GL_MACRO) {
      // Record 0 is one whole Poseidon-Goldilocks permutation, cooperatively: 12 threads fetch the state, warp 0 evaluates
      // the permutation natively (lane k owns element k) leaving the 1992 hint / product integers in shared memory, then all
      // threads convert them to Montgomery form and store them to their wires and to the ring. The permutation is a ~45 us
      // dependency chain on ONE warp: the chunk's other records - instructions of the same level, independent of it - are
      // executed by warps 1..7 meanwhile.
      const uint32_t* rec = ch + ch[4];
      const uint32_t nout = rec[0] >> 8;
      const uint32_t* t = rec + 7;
      const uint32_t* outs = t + 24;
      const uint32_t slot0 = rec[5];
      if (threadIdx.x < 12) {
        const LeRef A{t, t + 1, 12, 2, true, ring};
        uint64_t x[4];
        to_u64x4(eval_term(c, W, A, threadIdx.x), x);
        if ((x[1] | x[2] | x[3]) != 0 || x[0] >= gl::P) atomicCAS(e, 0, ERR_MULADD);  // goldilocks/base.go:228-232
        glm_io[threadIdx.x] = x[0];
      }
      __syncthreads();
      lap(1);
      if (threadIdx.x < 32) {
        glm::trace_warp(threadIdx.x < 12 ? glm_io[threadIdx.x] : 0ull, glm_io + 12, c.gl_tables, [&](uint32_t slot, const glm::U192& v) {
          glm_trace[3 * slot] = v.l[0];
          glm_trace[3 * slot + 1] = v.l[1];
          glm_trace[3 * slot + 2] = v.l[2];
        });
      } else {
        for (uint32_t i = 1 + (threadIdx.x - 32); i < n_instr; i += blockDim.x - 32) run_record(i);
      }
      if (blockDim.x == 32)  // (GPW_SPINE_THREADS=32 experiments: no helper warps)
        for (uint32_t i = 1 + threadIdx.x; i < n_instr; i += 32) run_record(i);
      __syncthreads();
      lap(2);
      for (uint32_t k = threadIdx.x; k < nout; k += blockDim.x) {
        const Fr v = fr_from_u192(glm::U192{{glm_trace[3 * k], glm_trace[3 * k + 1], glm_trace[3 * k + 2]}});
        st_w(W + outs[k], v);
        st_w(ring + ((slot0 + k) & (RING_SLOTS - 1)), v);
      }
      pc[5]++;
    } else {
      for (uint32_t i = threadIdx.x; i < n_instr; i += blockDim.x) run_record(i);
    }
    src = next_src;
    words = next_words;
    cur ^= 1;
    __syncthreads();  // everyone is done reading chunk `cur^1`... (now the old buffer) before it is overwritten
    lap(ch[3] == CHUNK_FLAG_GL_MACRO ? 3 : 4);
    pc[6]++;
  }
  if (prof && threadIdx.x == 0 && blockIdx.x == 0)
    for (int k = 0; k < 7; k++) atomicAdd(prof + k, pc[k]);
}

// one CTA per proof walks levels [lo, hi)
__global__ void __launch_bounds__(NARROW_THREADS)
    k_tape_narrow(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, int* __restrict__ err, uint32_t* __restrict__ hist,
                  uint32_t lo, uint32_t hi) {
  Fr* W = wires + (size_t)blockIdx.x * wire_stride;
  int* e = err + blockIdx.x;
  uint32_t* h = hist + (size_t)blockIdx.x * 65536;
  for (uint32_t lvl = lo; lvl < hi; lvl++) {
    const uint32_t s = c.level_off[lvl], t = c.level_off[lvl + 1];
    for (uint32_t i = s + threadIdx.x; i < t; i += blockDim.x) exec_instr(c, W, c.instr[i], e, h);
    __syncthreads();
  }
}

// grid.y = proof; instructions [s, t) of one level
__global__ void __launch_bounds__(128)
    k_tape_wide(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, int* __restrict__ err, uint32_t* __restrict__ hist,
                uint32_t s, uint32_t t) {
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  for (uint32_t i = s + blockIdx.x * blockDim.x + threadIdx.x; i < t; i += gridDim.x * blockDim.x)
    exec_instr(c, W, c.instr[i], err + blockIdx.y, hist + (size_t)blockIdx.y * 65536);
}

// grid.y = proof; the Poseidon-BN254 macro instructions [s, t) of one level, FOUR lanes per permutation (lane q evaluates
// input expression q and owns state element q, poseidon_bn254_trace4), 8 permutations per one-warp CTA so that the
// ~20..170 permutations of a Merkle-path level spread over as many SMs as possible: the level is a dependency chain,
// what counts is the pace of a single warp.
__global__ void __launch_bounds__(32)
    k_tape_poseidon4(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, uint32_t s, uint32_t t) {
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  const uint32_t lane = threadIdx.x, q = lane & 3u;
  const uint32_t i = s + blockIdx.x * 8u + (lane >> 2);
  const bool valid = i < t;
  const DInstr in = c.instr[valid ? i : s];
  const LeRef R = csr_ref(c, in.le[q]);
  Fr x = eval_ref(c, W, R);
  const uint32_t const_mask = (__ballot_sync(0xffffffffu, ref_is_const(R)) >> (lane & ~3u)) & 0xfu;
  const Bn254PoseidonTables T{c.bn_tables, c.bn_tables + 88, c.bn_tables + 88 + 392, c.bn_tables + 88 + 392 + 16};
  Fr* out = W + in.out;
  poseidon_bn254_trace4(x, const_mask, T, [&](uint32_t idx, const Fr& v) {
    if (valid) st_w(out + idx, v);
  });
}

// The Fr inversions of a wide level - gnark's IsZero hint (one per range check) and the 2.5 M divisions of the
// log-derivative argument - with Montgomery's trick: a thread owns INV_G instructions (strided by the grid so that
// neighbouring lanes touch neighbouring wires), multiplies their denominators up, inverts the product once (Fermat,
// ~380 multiplies) and peels the individual inverses off on the way back: ~15 multiplies per inversion instead of ~380.
// Zero denominators (IsZero of 0 -> 0; a division by zero is reported and yields 0 like inv(0) = 0 did) sit out.
constexpr int INV_G = 32;
__global__ void __launch_bounds__(128)
    k_tape_wide_inv(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, int* __restrict__ err, uint32_t s, uint32_t t) {
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  const uint32_t first = s + blockIdx.x * blockDim.x + threadIdx.x;
  Fr pre[INV_G], den[INV_G];
  uint32_t zero_mask = 0;
  int cnt = 0;
  Fr run = Fr::one();
#pragma unroll 1
  for (int j = 0; j < INV_G; j++) {
    const uint64_t i = (uint64_t)first + (uint64_t)j * nthreads;
    if (i >= t) break;
    const DInstr in = c.instr[i];
    const bool is_div = (in.op_nout & 0xffu) == fe::OP_DIV;
    Fr d = eval_le(c, W, is_div ? in.le[1] : in.le[0]);
    if (d.is_zero()) {
      if (is_div) atomicCAS(err + blockIdx.y, 0, ERR_DIV0);
      zero_mask |= 1u << j;
      d = Fr::one();
    }
    den[j] = d;
    run = mul(run, d);
    pre[j] = run;
    cnt = j + 1;
  }
  if (cnt == 0) return;
  Fr iv = inv(run);
#pragma unroll 1
  for (int j = cnt - 1; j >= 0; j--) {
    const uint32_t i = first + (uint32_t)j * nthreads;
    const DInstr in = c.instr[i];
    Fr r = j ? mul(iv, pre[j - 1]) : iv;
    iv = mul(iv, den[j]);
    if ((in.op_nout & 0xffu) == fe::OP_DIV) r = mul(eval_le(c, W, in.le[0]), r);
    if ((zero_mask >> j) & 1u) r = Fr::zero();
    st_w(W + in.out, r);
  }
}

__global__ void k_counts_to_wires(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, const uint32_t* __restrict__ hist) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 65536) return;
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  st_w(W + c.count_start + i, from_u64(hist[(size_t)blockIdx.y * 65536 + i]));
}

// inputs: canonical 4 x u64 per input (public then secret) -> wires 1.. in Montgomery form; wire 0 = 1
__global__ void k_set_inputs(Fr* __restrict__ wires, size_t wire_stride, const uint64_t* __restrict__ inputs, uint32_t n_inputs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  if (i == 0) st_w(W, Fr::one());
  if (i >= n_inputs) return;
  const uint64_t* src = inputs + ((size_t)blockIdx.y * n_inputs + i) * 4;
  uint64_t x[4] = {src[0], src[1], src[2], src[3]};
  st_w(W + 1 + i, from_u64x4(x));
}

// Long linear expressions (the two sides of the log-derivative identity: 65 536 and ~2.5 M terms). grid = (expression,
// slice): every CTA sums a strided slice of the terms (registers, then a shared-memory tree) into part[expression][slice];
// k_sum_long_les adds the LONG_SLICES partial sums. (One CTA per expression took 2.4 ms on one SM for the 2.5 M-term row.)
constexpr uint32_t LONG_SLICES = 128;
__global__ void __launch_bounds__(256) k_eval_long_les(DevCircuit c, const Fr* __restrict__ W, Fr* __restrict__ part) {
  __shared__ uint4 sm_raw[256 * 2];
  Fr* sm = reinterpret_cast<Fr*>(sm_raw);
  const uint32_t le = c.long_le[blockIdx.x];
  const uint32_t s = c.le_off[le], e = c.le_off[le + 1];
  Fr acc = Fr::zero();
  for (uint32_t k = s + blockIdx.y * blockDim.x + threadIdx.x; k < e; k += gridDim.y * blockDim.x) {
    const uint32_t cid = c.le_coeff[k];
    const Fr v = ld_w(W + c.le_wire[k]);
    if (cid == 0) acc = add(acc, v);
    else if (cid == 1) acc = sub(acc, v);
    else acc = add(acc, mul(ld_w(c.coeffs + cid), v));
  }
  st_w(sm + threadIdx.x, acc);
  __syncthreads();
  for (uint32_t d = blockDim.x >> 1; d >= 1; d >>= 1) {
    if (threadIdx.x < d) {
      acc = add(acc, ld_w(sm + threadIdx.x + d));
      st_w(sm + threadIdx.x, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_w(part + (size_t)blockIdx.x * gridDim.y + blockIdx.y, acc);
}

__global__ void __launch_bounds__(LONG_SLICES) k_sum_long_les(DevCircuit c, const Fr* __restrict__ part) {
  __shared__ uint4 sm_raw[LONG_SLICES * 2];
  Fr* sm = reinterpret_cast<Fr*>(sm_raw);
  Fr acc = ld_w(part + (size_t)blockIdx.x * LONG_SLICES + threadIdx.x);
  st_w(sm + threadIdx.x, acc);
  __syncthreads();
  for (uint32_t d = LONG_SLICES >> 1; d >= 1; d >>= 1) {
    if (threadIdx.x < d) {
      acc = add(acc, ld_w(sm + threadIdx.x + d));
      st_w(sm + threadIdx.x, acc);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) st_w(c.long_val + blockIdx.x, acc);
}

// a = L.w, b = R.w, c = O.w for every constraint (zero-padded to the FFT domain by the caller's memset);
// counts rows with a*b != c.
__global__ void __launch_bounds__(128)
    k_r1cs_eval(DevCircuit c, const Fr* __restrict__ W, Fr* __restrict__ a, Fr* __restrict__ b, Fr* __restrict__ cc,
                unsigned long long* __restrict__ n_bad, unsigned long long* __restrict__ first_bad) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n_cons) return;
  Fr v[3];
#pragma unroll 1
  for (int j = 0; j < 3; j++) {
    const uint32_t le = c.cons[3 * k + j];
    bool is_long = false;
    for (uint32_t q = 0; q < c.n_long; q++)
      if (c.long_le[q] == le) {
        v[j] = ld_w(c.long_val + q);
        is_long = true;
      }
    if (!is_long) v[j] = eval_le(c, W, le);
  }
  const Fr &l = v[0], &r = v[1], &o = v[2];
  if (a) {
    st_w(a + k, l);
    st_w(b + k, r);
    st_w(cc + k, o);
  }
  if (mul(l, r) != o) {
    atomicAdd(n_bad, 1ull);
    atomicMin(first_bad, (unsigned long long)k);
  }
}

}  // namespace gpw

using namespace gpw;

struct gpw_circuit {
  gpw_ctx* ctx = nullptr;
  fe::API api;
  gadgets::CommonCircuitData cd;
  bool is_verifier = false;
  DevCircuit dc{};
  std::vector<Segment> plan;
  std::vector<void*> dev_allocs;
  const uint32_t* stream_dev = nullptr;
  double ring_hit_rate = 0;
  uint32_t n_inputs = 0;
  float solve_ms = 0;
  // leading secret-input values compiled in as constants (gpw_circuit_compile_verifier_bound): parse_inputs checks the
  // documents against them and leaves them out of the input vector
  std::vector<std::array<uint64_t, 4>> baked;
  std::string common_json;  // kept for the compile cache (gpw_circuit_save)
};

template <class T>
static int upload(gpw_circuit* c, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = v.size() * sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return GPW_ENOMEM;
  }
  c->dev_allocs.push_back(p);
  if (bytes) GPW_CUDA(cudaMemcpy(p, v.data(), bytes, cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return GPW_OK;
}

static int finish_compile(gpw_circuit* c) {
  fe::API& api = c->api;
  api.ScheduleSpineAndTail();
  const auto& tape = api.Tape();
  const uint32_t L = api.NumLevels();
  // sort by (level, op)
  std::vector<uint32_t> order(tape.size());
  for (uint32_t i = 0; i < tape.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    if (tape[x].level != tape[y].level) return tape[x].level < tape[y].level;
    return tape[x].op < tape[y].op;
  });
  std::vector<DInstr> di;
  di.reserve(tape.size());
  std::vector<uint32_t> level_off(L + 1, 0);
  uint32_t count_level = 0xffffffffu, commit_level = 0xffffffffu;
  for (uint32_t idx : order) {
    const fe::Instr& in = tape[idx];
    if (in.op == fe::OP_COUNT) {
      count_level = in.level;
      continue;
    }
    if (in.op == fe::OP_COMMIT) {
      commit_level = in.level;
      continue;
    }
    // (a macro with scattered outputs carries the offset of its output list instead of a first wire)
    di.push_back({(uint32_t)in.op | (in.nout << 8), in.outs_off != NO_LE ? in.outs_off : in.out, {in.le[0], in.le[1], in.le[2], in.le3}});
    level_off[in.level + 1]++;
  }
  for (uint32_t l = 0; l < L; l++) level_off[l + 1] += level_off[l];
  // execution plan
  c->plan.clear();
  static const bool poseidon4 = !getenv("GPW_NO_POSEIDON4");  // (debugging aid: everything on the one-thread-per-permutation path)
  uint32_t run_lo = 0, run_skip_s = 0, run_skip_t = 0;
  auto flush = [&](uint32_t upto) {
    if (upto > run_lo) {
      Segment n{SEG_NARROW, run_lo, upto};
      n.skip_s = run_skip_s;
      n.skip_t = run_skip_t;
      c->plan.push_back(n);
    }
    run_skip_s = run_skip_t = 0;
  };
  // instructions of a level are sorted by op: 0 = generic, 1 = Fr inversion (batched-inverse kernel), 2 = Poseidon-BN254 macro
  auto op_class = [&](uint32_t i) -> int {
    const uint32_t op = di[i].op_nout & 0xffu;
    return (op == fe::OP_INVZERO || op == fe::OP_DIV) ? 1 : op == fe::OP_POSEIDON_BN254 ? 2 : 0;
  };
  for (uint32_t l = 0; l < L; l++) {
    const uint32_t cnt = level_off[l + 1] - level_off[l];
    const bool special = (l == count_level || l == commit_level);
    if (cnt >= wide_threshold() || special) {
      flush(l);
      if (cnt) {
        Segment w{SEG_WIDE, l, l + 1};
        uint32_t i = level_off[l];
        while (i < level_off[l + 1]) {
          uint32_t j = i;
          const int cl = op_class(i);
          while (j < level_off[l + 1] && op_class(j) == cl) j++;
          Segment::Part part{i, j, cl == 1 && (j - i) >= 1024};
          part.poseidon = poseidon4 && cl == 2 && (j - i) >= POSEIDON4_MIN;
          w.parts.push_back(part);
          i = j;
        }
        // merge neighbouring generic parts (short inversion / permutation runs stay on the generic path)
        std::vector<Segment::Part> merged;
        for (const auto& p : w.parts) {
          const bool plain = !p.inv && !p.poseidon;
          if (!merged.empty() && plain && !merged.back().inv && !merged.back().poseidon) merged.back().t = p.t;
          else merged.push_back(p);
        }
        w.parts = merged;
        c->plan.push_back(w);
      }
      if (l == count_level) c->plan.push_back({SEG_COUNT, l, l + 1});
      if (l == commit_level) c->plan.push_back({SEG_COMMIT, l, l + 1});
      run_lo = l + 1;
    } else if (poseidon4) {
      // a narrow level with enough Poseidon-BN254 macros: the spine stops in front of it, the four-lane grid kernel runs the
      // permutations, and the spine resumes AT this level without them (instructions of one level are independent)
      uint32_t ps = level_off[l], pt;
      while (ps < level_off[l + 1] && op_class(ps) != 2) ps++;
      pt = ps;
      while (pt < level_off[l + 1] && op_class(pt) == 2) pt++;
      if (pt - ps >= POSEIDON4_MIN) {
        flush(l);
        Segment g{SEG_POSEIDON, l, l + 1};
        g.ps = ps;
        g.pt = pt;
        c->plan.push_back(g);
        run_lo = l;
        run_skip_s = ps;
        run_skip_t = pt;
      }
    }
  }
  flush(L);
  // staged stream for the narrow segments (see k_tape_staged)
  {
    const auto& off = api.LeOffsets();
    const auto& lw = api.LeWires();
    const auto& lc = api.LeCoeffIds();
    std::vector<uint32_t> stream;
    auto le_words = [&](uint32_t le) -> uint32_t { return le == NO_LE ? 0 : 2 * (off[le + 1] - off[le]); };
    const bool dbg = getenv("GPW_DEBUG_TAPE") != nullptr;
    uint64_t dbg_thread[16][3] = {}, dbg_warp[3] = {0, 0, 0}, dbg_warps = 0, dbg_chunks = 0;
    std::vector<uint8_t> coeff_small(api.Coeffs().size(), 0);
    for (size_t ci = 0; ci < api.Coeffs().size(); ci++) {
      uint64_t l4[4];
      fe::fr_to_limbs(api.Coeffs()[ci], l4);
      coeff_small[ci] = (l4[1] | l4[2] | l4[3]) == 0;
    }
    std::vector<uint64_t> wire_seq(api.NumWires(), 0);
    std::vector<uint32_t> wire_seg(api.NumWires(), 0xffffffffu);
    uint64_t ring_seq = 0, ring_hits = 0, ring_total = 0;
    uint32_t seg_id = 0;
    for (auto& seg : c->plan) {
      if (seg.kind != SEG_NARROW) continue;
      seg_id++;
      seg.stream_off = (uint32_t)stream.size();
      seg.first_words = 0;
      size_t prev_hdr = (size_t)-1;
      for (uint32_t l = seg.lo; l < seg.hi; l++) {
        // the level's instructions in chunk order: a Poseidon-Goldilocks macro FIRST (it opens a chunk and the level's other
        // instructions fill the rest of it - warp 0 runs the permutation while the other warps execute them), then the rest.
        // The first level of a segment may leave a run of instructions (skip_s, skip_t) to a preceding SEG_POSEIDON.
        const bool cut = l == seg.lo && seg.skip_t > seg.skip_s;
        std::vector<DInstr> dl;
        dl.reserve(level_off[l + 1] - level_off[l]);
        for (int pass = 0; pass < 2; pass++)
          for (uint32_t k = level_off[l]; k < level_off[l + 1]; k++) {
            if (cut && k >= seg.skip_s && k < seg.skip_t) continue;
            const bool glm = (di[k].op_nout & 0xffu) == fe::OP_POSEIDON_GL;
            if (glm == (pass == 0)) dl.push_back(di[k]);
          }
        uint32_t i = 0;
        const uint32_t end = (uint32_t)dl.size();
        while (i < end) {
          // greedily take instructions [i, j) that fit one chunk
          uint32_t j = i, words = 4, outs_sum = 0;
          const auto& mouts = api.MacroOuts();
          auto is_gl_macro = [&](uint32_t k) { return (dl[k].op_nout & 0xffu) == fe::OP_POSEIDON_GL; };
          while (j < end) {
            if (is_gl_macro(j) && j > i) break;  // a Poseidon-Goldilocks macro opens a chunk
            uint32_t rec = 7 + le_words(dl[j].le[0]) + le_words(dl[j].le[1]) + le_words(dl[j].le[2]) + le_words(dl[j].le[3]);
            if (is_gl_macro(j)) rec += dl[j].op_nout >> 8;
            if ((dl[j].op_nout >> 8) >= RING_SLOTS) {
              set_error("tape instruction with %u outputs exceeds the ring", dl[j].op_nout >> 8);
              return GPW_EINVAL;
            }
            if (rec + 5 > CHUNK_MAX_WORDS) {
              set_error("tape instruction too large for a staged chunk (%u words)", rec);
              return GPW_EINVAL;
            }
            if (words + 1 + rec > CHUNK_MAX_WORDS - 4) break;
            // the records of a chunk write their ring slots in no particular order: together they must not wrap around the
            // ring (two wires of one chunk on one slot); a macro's own outputs are stored after the others (below)
            if (!is_gl_macro(j)) {
              if (outs_sum + (dl[j].op_nout >> 8) > RING_SLOTS) break;
              outs_sum += dl[j].op_nout >> 8;
            }
            words += 1 + rec;
            j++;
          }
          const size_t base = stream.size();
          const uint32_t n = j - i;
          stream.resize(base + 4 + n, 0);
          stream[base] = n;
          stream[base + 2] = (j == end) ? 1u : 0u;
          stream[base + 3] = is_gl_macro(i) ? CHUNK_FLAG_GL_MACRO : 0u;
          // ring bookkeeping: every wire produced on the spine gets the next slot of the shared-memory ring; a later
          // operand reads the ring instead of HBM if its slot cannot have been overwritten before the END of the
          // consuming chunk (the chunk's own outputs are written concurrently with its reads)
          uint64_t seq_end = ring_seq;
          for (uint32_t k = i; k < j; k++) seq_end += dl[k].op_nout >> 8;
          if (dbg) {  // warp-level cost model: lanes of a warp run in lock step, the slowest lane sets the pace
            for (uint32_t w0 = i; w0 < j; w0 += 32) {
              uint32_t worst[3] = {0, 0, 0};
              for (uint32_t k = w0; k < std::min(j, w0 + 32); k++) {
                uint32_t cnt[3] = {0, 0, 0};
                for (int t = 0; t < 4; t++) {
                  uint32_t le = dl[k].le[t];
                  if (le == NO_LE) continue;
                  for (uint32_t q = off[le]; q < off[le + 1]; q++) {
                    if (lc[q] <= 1) continue;
                    cnt[lw[q] == 0 ? 0 : (coeff_small[lc[q]] ? 1 : 2)]++;
                  }
                }
                for (int t = 0; t < 3; t++) worst[t] = std::max(worst[t], cnt[t]);
                dbg_thread[dl[k].op_nout & 0xff][0] += cnt[0];
                dbg_thread[dl[k].op_nout & 0xff][1] += cnt[1];
                dbg_thread[dl[k].op_nout & 0xff][2] += cnt[2];
              }
              for (int t = 0; t < 3; t++) dbg_warp[t] += worst[t];
              dbg_warps++;
            }
            dbg_chunks++;
          }
          // ring slots in the order the chunk WRITES them: a macro (record 0) stores its outputs after the chunk's other
          // records have run (k_tape_staged), so it takes the last slots of the chunk - a slot always holds the newest wire
          std::vector<uint64_t> seq_start(j - i);
          {
            uint64_t rs = ring_seq;
            const bool macro_first = is_gl_macro(i);
            for (uint32_t k = i + (macro_first ? 1u : 0u); k < j; k++) {
              seq_start[k - i] = rs;
              rs += dl[k].op_nout >> 8;
            }
            if (macro_first) seq_start[0] = rs;
          }
          for (uint32_t k = i; k < j; k++) {
            stream[base + 4 + (k - i)] = (uint32_t)(stream.size() - base);
            const DInstr& in = dl[k];
            uint64_t my_seq = seq_start[k - i];
            stream.push_back(in.op_nout);
            stream.push_back(in.out);
            const bool same_ab = (in.op_nout & 0xff) == fe::OP_MUL && in.le[1] != NO_LE && in.le[1] == in.le[0];
            for (int t = 0; t < 3; t++) {
              uint32_t le = in.le[t];
              if (t == 1 && same_ab) stream.push_back(0xC0000000u);
              else stream.push_back(le == NO_LE ? 0u : ((off[le + 1] - off[le]) | 0x80000000u));
            }
            stream.push_back((uint32_t)(my_seq % RING_SLOTS));
            stream.push_back(in.le[3] == NO_LE ? 0u : ((off[in.le[3] + 1] - off[in.le[3]]) | 0x80000000u));
            const bool glm_instr = (in.op_nout & 0xffu) == fe::OP_POSEIDON_GL;
            for (uint32_t o = 0; o < (in.op_nout >> 8); o++) {
              const uint32_t ow = glm_instr ? mouts[in.out + o] : in.out + o;
              wire_seq[ow] = my_seq++;
              wire_seg[ow] = seg_id;
            }
            for (int t = 0; t < 4; t++) {
              uint32_t le = in.le[t];
              if (le == NO_LE || (t == 1 && same_ab)) continue;
              for (uint32_t q = off[le]; q < off[le + 1]; q++) {
                const uint32_t w = lw[q];
                const bool in_ring = wire_seg[w] == seg_id && seq_end - wire_seq[w] <= RING_SLOTS;
                stream.push_back(in_ring ? (RING_FLAG | (uint32_t)(wire_seq[w] % RING_SLOTS)) : w);
                stream.push_back(lc[q]);
                ring_hits += in_ring;
                ring_total++;
              }
            }
            if (glm_instr)
              for (uint32_t o = 0; o < (in.op_nout >> 8); o++) stream.push_back(mouts[in.out + o]);
          }
          ring_seq = seq_end;
          while ((stream.size() - base) % 4) stream.push_back(0);
          const uint32_t cw = (uint32_t)(stream.size() - base);
          if (prev_hdr == (size_t)-1) seg.first_words = cw;
          else stream[prev_hdr + 1] = cw;
          prev_hdr = base;
          i = j;
        }
      }
    }
    stream.resize(stream.size() + 8, 0);
    c->ring_hit_rate = ring_total ? (double)ring_hits / (double)ring_total : 0.0;
    if (dbg) {
      fprintf(stderr, "[gpw tape] narrow stream: %llu chunks, %llu warp-passes, ring hit rate %.3f\n",
              (unsigned long long)dbg_chunks, (unsigned long long)dbg_warps, c->ring_hit_rate);
      fprintf(stderr, "[gpw tape] warp-level coefficient muls: const(wire0) %llu, small %llu, full %llu\n",
              (unsigned long long)dbg_warp[0], (unsigned long long)dbg_warp[1], (unsigned long long)dbg_warp[2]);
      for (int op = 0; op < 9; op++)
        fprintf(stderr, "[gpw tape] op %d thread-level coefficient terms: const %llu small %llu full %llu\n", op,
                (unsigned long long)dbg_thread[op][0], (unsigned long long)dbg_thread[op][1], (unsigned long long)dbg_thread[op][2]);
    }
    GPW_TRY(upload(c, stream, &c->stream_dev));
    GPW_CUDA(cudaFuncSetAttribute(k_tape_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SPINE_EXCLUSIVE_SMEM_BYTES));
  }
  DevCircuit& dc = c->dc;
  GPW_TRY(upload(c, di, &dc.instr));
  GPW_TRY(upload(c, level_off, &dc.level_off));
  GPW_TRY(upload(c, api.LeOffsets(), &dc.le_off));
  GPW_TRY(upload(c, api.LeWires(), &dc.le_wire));
  GPW_TRY(upload(c, api.LeCoeffIds(), &dc.le_coeff));
  GPW_TRY(upload(c, api.Coeffs(), &dc.coeffs));
  GPW_TRY(upload(c, api.Constraints(), &dc.cons));
  dc.n_wires = api.NumWires();
  dc.n_cons = (uint32_t)api.NumConstraints();
  dc.n_levels = L;
  dc.limb_start = api.LimbWireStart();
  dc.n_limbs = api.NumLimbWires();
  dc.count_start = api.CountWireStart();
  dc.commit_wire = api.CommitWire();
  {
    std::vector<Fr> tb(88 + 392 + 16 + 16);
    memcpy(tb.data(), GPW_BN_C_MONT, 88 * 32);
    memcpy(tb.data() + 88, GPW_BN_S_MONT, 392 * 32);
    memcpy(tb.data() + 88 + 392, GPW_BN_M_MONT, 16 * 32);
    memcpy(tb.data() + 88 + 392 + 16, GPW_BN_P_MONT, 16 * 32);
    GPW_TRY(upload(c, tb, &dc.bn_tables));
  }
  {
    std::vector<uint64_t> gt(glm::T_TOTAL);
    memcpy(gt.data() + glm::T_RC, GPW_GL_ALL_ROUND_CONSTANTS, sizeof(GPW_GL_ALL_ROUND_CONSTANTS));
    memcpy(gt.data() + glm::T_CIRC, GPW_GL_MDS_CIRC, sizeof(GPW_GL_MDS_CIRC));
    memcpy(gt.data() + glm::T_DIAG, GPW_GL_MDS_DIAG, sizeof(GPW_GL_MDS_DIAG));
    memcpy(gt.data() + glm::T_FIRST, GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT, sizeof(GPW_GL_FAST_PARTIAL_FIRST_ROUND_CONSTANT));
    memcpy(gt.data() + glm::T_PRC, GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_CONSTANTS));
    memcpy(gt.data() + glm::T_VS, GPW_GL_FAST_PARTIAL_ROUND_VS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_VS));
    memcpy(gt.data() + glm::T_WHATS, GPW_GL_FAST_PARTIAL_ROUND_W_HATS, sizeof(GPW_GL_FAST_PARTIAL_ROUND_W_HATS));
    memcpy(gt.data() + glm::T_INIT, GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX, sizeof(GPW_GL_FAST_PARTIAL_ROUND_INITIAL_MATRIX));
    static_assert(GPW_GL_MDS0TO0 == glm::MDS0TO0, "MDS0TO0");
    GPW_TRY(upload(c, gt, &dc.gl_tables));
    std::vector<uint32_t> mo = api.MacroOuts();
    if (mo.empty()) mo.push_back(0);
    GPW_TRY(upload(c, mo, &dc.macro_outs));
  }
  dc.n_long = 0;
  {
    std::vector<uint8_t> used(api.LeOffsets().size(), 0);
    for (uint32_t le : api.Constraints()) used[le] = 1;
    const auto& off = api.LeOffsets();
    for (uint32_t le = 0; le + 1 < off.size(); le++)
      if (used[le] && off[le + 1] - off[le] > LONG_LE_TERMS) {
        if (dc.n_long >= 8) {
          set_error("circuit has more than 8 very long linear expressions");
          return GPW_EINVAL;
        }
        dc.long_le[dc.n_long++] = le;
      }
    std::vector<Fr> zeros(8, Fr::zero());
    const Fr* lv;
    GPW_TRY(upload(c, zeros, &lv));
    dc.long_val = const_cast<Fr*>(lv);
  }
  c->n_inputs = api.NumPublic() + api.NumSecret();
  return GPW_OK;
}

extern "C" void gpw_circuit_free(gpw_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  for (void* p : c->dev_allocs) cudaFree(p);
  delete c;
}

extern "C" int gpw_circuit_compile_verifier_bound(gpw_ctx* ctx, const char* common_circuit_data_json, const char* verifier_only_json,
                                                  const char* proof_json, gpw_circuit** out);
extern "C" int gpw_circuit_compile_verifier(gpw_ctx* ctx, const char* common_circuit_data_json, gpw_circuit** out) {
  return gpw_circuit_compile_verifier_bound(ctx, common_circuit_data_json, nullptr, nullptr, out);
}

// verifier_only_json != NULL: VerifierOnlyCircuitData (constants_sigmas_cap, circuit_digest) are compile-time constants, as
// the reference's `gnark:"-"` tag makes them (verifier/util.go:13) - the compiled circuit then accepts only proofs of that
// inner circuit. proof_json != NULL as well: the proof itself is baked in too, which is literally benchmark.go:33-55.
extern "C" int gpw_circuit_compile_verifier_bound(gpw_ctx* ctx, const char* common_circuit_data_json, const char* verifier_only_json,
                                                  const char* proof_json, gpw_circuit** out) {
  if (!ctx || !common_circuit_data_json || !out || (proof_json && !verifier_only_json)) {
    set_error("circuit_compile: null argument (a baked proof needs the verifier-only data too)");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  gpw_circuit* c = new gpw_circuit();
  c->ctx = ctx;
  try {
    c->cd = gadgets::ReadCommonCircuitData(common_circuit_data_json);
    c->common_json = common_circuit_data_json;
    if (proof_json) c->baked = gadgets::ParseProofInputs(c->cd, proof_json, verifier_only_json).sec;
    else if (verifier_only_json) c->baked = gadgets::ParseVerifierOnly(c->cd, verifier_only_json);
    gadgets::DefineVerifierCircuit(&c->api, c->cd, c->baked.empty() ? nullptr : &c->baked);
    c->is_verifier = true;
  } catch (const std::exception& e) {
    set_error("circuit_compile: %s", e.what());
    delete c;
    return GPW_EINVAL;
  }
  int rc = finish_compile(c);
  if (rc != GPW_OK) {
    gpw_circuit_free(c);
    return rc;
  }
  *out = c;
  return GPW_OK;
}

// info: [wires, public, secret, constraints, instructions, levels, limb_wires, limb_start, count_start, commit_wire,
//        n_narrow_segments, n_wide_segments, muladd, reduce, glinv, split]
extern "C" int gpw_circuit_info(const gpw_circuit* c, uint64_t* info16) {
  if (!c || !info16) return GPW_EINVAL;
  uint64_t nn = 0, nw = 0;
  for (const auto& s : c->plan) {
    if (s.kind == SEG_NARROW) nn++;
    if (s.kind == SEG_WIDE) nw++;
  }
  const auto& k = c->api.Counts();
  uint64_t v[16] = {c->api.NumWires(), c->api.NumPublic(), c->api.NumSecret(), c->api.NumConstraints(), c->api.Tape().size(),
                    c->api.NumLevels(), c->api.NumLimbWires(), c->api.LimbWireStart(), c->api.CountWireStart(),
                    c->api.CommitWire(), nn, nw, k.muladd, k.reduce, k.glinv, k.split};
  memcpy(info16, v, sizeof(v));
  return GPW_OK;
}

static const char* err_name(int e) {
  switch (e) {
    case ERR_MULADD: return "MulAddHint: operand is not in the field (goldilocks/base.go:228-232)";
    case ERR_GLINV: return "InverseHint: input is not in the field (goldilocks/base.go:322-324)";
    case ERR_SPLIT: return "SplitLimbsHint: input is not in the field (goldilocks/base.go:347-349)";
    case ERR_BITS: return "ToBinary: value does not fit in the requested number of bits";
    case ERR_DECOMP: return "range check: value exceeds its bit width";
    case ERR_DIV0: return "log-derivative argument: division by zero (challenge collides with a table entry)";
    default: return "unknown";
  }
}

// Runs plan segments [seg_lo, seg_hi) for n_proofs proofs.
static int run_segments(gpw_circuit* c, gpw_ctx* ctx, Fr* wires, size_t stride, int n_proofs, int* err, uint32_t* hist,
                        size_t seg_lo, size_t seg_hi) {
  cudaStream_t st = ctx->stream;
  for (size_t si = seg_lo; si < seg_hi; si++) {
    const Segment& s = c->plan[si];
    if (s.kind == SEG_NARROW) {
      if (!s.first_words) continue;
      if (getenv("GPW_SOLVER_UNSTAGED")) {  // debugging aid: the simple per-level walker over the CSR arrays
        k_tape_narrow<<<n_proofs, spine_threads(), 0, st>>>(c->dc, wires, stride, err, hist, s.lo, s.hi);
      } else {
        unsigned long long* prof = nullptr;
        if (getenv("GPW_SPINE_PROFILE")) {
          GPW_TRY(ctx->get_scratch("solve.prof", 7 * 8, (void**)&prof));
          if (si == seg_lo) GPW_CUDA(cudaMemsetAsync(prof, 0, 7 * 8, st));
        }
        k_tape_staged<<<n_proofs, spine_threads(), spine_smem_bytes(), st>>>(c->dc, c->stream_dev + s.stream_off, s.first_words, wires, stride, err,
                                                                           hist, prof);
      }
      GPW_CHECK_LAUNCH();
      ctx->launches++;
    } else if (s.kind == SEG_POSEIDON) {
      dim3 grid(div_up(s.pt - s.ps, 8), n_proofs);
      k_tape_poseidon4<<<grid, 32, 0, st>>>(c->dc, wires, stride, s.ps, s.pt);
      GPW_CHECK_LAUNCH();
      ctx->launches++;
    } else if (s.kind == SEG_WIDE) {
      for (const Segment::Part& p : s.parts) {
        if (p.poseidon) {
          dim3 grid(div_up(p.t - p.s, 8), n_proofs);
          k_tape_poseidon4<<<grid, 32, 0, st>>>(c->dc, wires, stride, p.s, p.t);
        } else if (p.inv) {
          dim3 grid(div_up(p.t - p.s, 128 * INV_G), n_proofs);
          k_tape_wide_inv<<<grid, 128, 0, st>>>(c->dc, wires, stride, err, p.s, p.t);
        } else {
          dim3 grid(std::min<int>(ctx->sm_count * 8, div_up(p.t - p.s, 128)), n_proofs);
          k_tape_wide<<<grid, 128, 0, st>>>(c->dc, wires, stride, err, hist, p.s, p.t);
        }
        GPW_CHECK_LAUNCH();
        ctx->launches++;
      }
    } else if (s.kind == SEG_COUNT) {
      dim3 grid(65536 / 256, n_proofs);
      k_counts_to_wires<<<grid, 256, 0, st>>>(c->dc, wires, stride, hist);
      GPW_CHECK_LAUNCH();
      ctx->launches++;
    }
  }
  return GPW_OK;
}

static size_t commit_segment(const gpw_circuit* c) {
  for (size_t i = 0; i < c->plan.size(); i++)
    if (c->plan[i].kind == SEG_COMMIT) return i;
  return c->plan.size();
}

static void print_spine_profile(gpw_ctx* ctx) {
  if (!getenv("GPW_SPINE_PROFILE")) return;
  unsigned long long* prof = nullptr;
  if (ctx->get_scratch("solve.prof", 7 * 8, (void**)&prof) != GPW_OK) return;
  unsigned long long h[7];
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
  const double us = 1.0 / 1965.0;  // cycles -> microseconds at the B200's 1965 MHz
  fprintf(stderr, "[gpw spine] wait %.1f ms | macro: fetch %.1f, permutation %.1f, convert+store %.1f ms (%llu macros) | other chunks %.1f ms (%llu chunks)\n",
          h[0] * us / 1e3, h[1] * us / 1e3, h[2] * us / 1e3, h[3] * us / 1e3, h[5], h[4] * us / 1e3, h[6] - h[5]);
}

static int check_err(gpw_ctx* ctx, int* err_dev, int n_proofs) {
  if ((size_t)n_proofs * sizeof(int) > gpw_ctx::PIN_CAP) {
    set_error("witness_solve: too many proofs in one call");
    return GPW_EINVAL;
  }
  const int* e = (const int*)ctx->pin_take(n_proofs * sizeof(int));
  GPW_CUDA(cudaMemcpyAsync((void*)e, err_dev, n_proofs * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n_proofs; i++)
    if (e[i]) {
      set_error("witness solve failed for proof %d: %s", i, err_name(e[i]));
      // a hint refusing its input mirrors the reference's panic / error return; a value that does not fit its
      // decomposition means the constraint system is unsatisfiable for this assignment
      return (e[i] == ERR_BITS || e[i] == ERR_DECOMP || e[i] == ERR_DIV0) ? GPW_EUNSAT : GPW_EHINT;
    }
  return GPW_OK;
}

// The *_on variants run on `lane`, any context of the circuit's device: a context is a stream plus its scratch
// memory, so several proofs of one compiled circuit can be in flight side by side (wrap.cu runs one lane per host
// thread). The circuit itself is read-only after compilation.
//
// Phase 1: everything up to (not including) the commitment challenge. inputs_dev: n_proofs x n_inputs x 4 u64
// canonical (public then secret). wires_dev: n_proofs x wire_stride Fr.
extern "C" int gpw_witness_solve_phase1_on(gpw_circuit* c, gpw_ctx* lane, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev,
                                           size_t wire_stride) {
  if (!c || !lane || !inputs_dev || !wires_dev || n_proofs < 1 || wire_stride < c->dc.n_wires || lane->device != c->ctx->device) {
    set_error("witness_solve: bad argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = lane;
  GPW_CUDA(cudaSetDevice(ctx->device));
  int* err;
  uint32_t* hist;
  GPW_TRY(ctx->get_scratch("solve.err", (size_t)n_proofs * sizeof(int), (void**)&err));
  GPW_TRY(ctx->get_scratch("solve.hist", (size_t)n_proofs * 65536 * 4, (void**)&hist));
  GPW_CUDA(cudaMemsetAsync(err, 0, (size_t)n_proofs * sizeof(int), ctx->stream));
  GPW_CUDA(cudaMemsetAsync(hist, 0, (size_t)n_proofs * 65536 * 4, ctx->stream));
  dim3 grid(div_up(std::max<uint32_t>(c->n_inputs, 1), 256), n_proofs);
  k_set_inputs<<<grid, 256, 0, ctx->stream>>>((Fr*)wires_dev, wire_stride, (const uint64_t*)inputs_dev, c->n_inputs);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  GPW_TRY(run_segments(c, ctx, (Fr*)wires_dev, wire_stride, n_proofs, err, hist, 0, commit_segment(c)));
  const int rc = check_err(ctx, err, n_proofs);
  print_spine_profile(ctx);
  return rc;
}

extern "C" int gpw_witness_solve_phase1_dev(gpw_circuit* c, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride) {
  return gpw_witness_solve_phase1_on(c, c ? c->ctx : nullptr, inputs_dev, n_proofs, wires_dev, wire_stride);
}

// Phase 2: sets the commitment challenge (one canonical Fr per proof) and runs the rest of the tape.
extern "C" int gpw_witness_solve_phase2_on(gpw_circuit* c, gpw_ctx* lane, const uint64_t* challenges_canonical, int n_proofs,
                                           uint64_t wires_dev, size_t wire_stride) {
  if (!c || !lane || !wires_dev || n_proofs < 1 || lane->device != c->ctx->device) {
    set_error("witness_solve: bad argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = lane;
  GPW_CUDA(cudaSetDevice(ctx->device));
  size_t cs = commit_segment(c);
  if (cs == c->plan.size()) return GPW_OK;  // circuit has no commitment
  if (!challenges_canonical) {
    set_error("witness_solve: circuit needs a commitment challenge");
    return GPW_EINVAL;
  }
  int* err;
  uint32_t* hist;
  GPW_TRY(ctx->get_scratch("solve.err", (size_t)n_proofs * sizeof(int), (void**)&err));
  GPW_TRY(ctx->get_scratch("solve.hist", (size_t)n_proofs * 65536 * 4, (void**)&hist));
  GPW_CUDA(cudaMemsetAsync(err, 0, (size_t)n_proofs * sizeof(int), ctx->stream));  // phase 1 reported its own status
  for (int p = 0; p < n_proofs; p++) {
    Fr* x = (Fr*)ctx->pin_take(sizeof(Fr));
    *x = fe::fr_from_limbs(challenges_canonical + 4 * p);
    GPW_CUDA(cudaMemcpyAsync((Fr*)wires_dev + (size_t)p * wire_stride + c->dc.commit_wire, x, sizeof(Fr), cudaMemcpyHostToDevice,
                             ctx->stream));
  }
  GPW_TRY(run_segments(c, ctx, (Fr*)wires_dev, wire_stride, n_proofs, err, hist, cs + 1, c->plan.size()));
  return check_err(ctx, err, n_proofs);
}

extern "C" int gpw_witness_solve_phase2_dev(gpw_circuit* c, const uint64_t* challenges_canonical, int n_proofs, uint64_t wires_dev,
                                            size_t wire_stride) {
  return gpw_witness_solve_phase2_on(c, c ? c->ctx : nullptr, challenges_canonical, n_proofs, wires_dev, wire_stride);
}

// a, b, c evaluation vectors of one proof (may be 0 to only check). Returns GPW_EUNSAT if a constraint fails.
extern "C" int gpw_r1cs_eval_on(gpw_circuit* c, gpw_ctx* lane, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                                uint64_t* n_unsatisfied) {
  if (!c || !lane || !wires_dev || lane->device != c->ctx->device) {
    set_error("r1cs_eval: bad argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = lane;
  GPW_CUDA(cudaSetDevice(ctx->device));
  // [0..1]: unsatisfied count / first bad row; then the lane's values of the circuit's long linear expressions
  unsigned long long* bad;
  GPW_TRY(ctx->get_scratch("solve.bad", 64 + 8 * sizeof(Fr) + 8 * LONG_SLICES * sizeof(Fr), (void**)&bad));
  DevCircuit dc = c->dc;
  dc.long_val = reinterpret_cast<Fr*>(bad + 8);
  Fr* long_part = dc.long_val + 8;
  unsigned long long* init = (unsigned long long*)ctx->pin_take(16);
  init[0] = 0;
  init[1] = ~0ull;
  GPW_CUDA(cudaMemcpyAsync(bad, init, 16, cudaMemcpyHostToDevice, ctx->stream));
  if (dc.n_long) {
    k_eval_long_les<<<dim3(dc.n_long, LONG_SLICES), 256, 0, ctx->stream>>>(dc, (const Fr*)wires_dev, long_part);
    GPW_CHECK_LAUNCH();
    k_sum_long_les<<<dc.n_long, LONG_SLICES, 0, ctx->stream>>>(dc, long_part);
    GPW_CHECK_LAUNCH();
    ctx->launches += 2;
  }
  if (dc.n_cons) {  // (a circuit without constraints - a NoopGate on its own - is trivially satisfied)
    k_r1cs_eval<<<div_up(dc.n_cons, 128), 128, 0, ctx->stream>>>(dc, (const Fr*)wires_dev, (Fr*)a_dev, (Fr*)b_dev, (Fr*)c_dev, bad, bad + 1);
    GPW_CHECK_LAUNCH();
    ctx->launches++;
  }
  const unsigned long long* res = (const unsigned long long*)ctx->pin_take(16);
  GPW_CUDA(cudaMemcpyAsync((void*)res, bad, 16, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_unsatisfied) *n_unsatisfied = res[0];
  if (res[0]) {
    set_error("%llu constraints unsatisfied (first: #%llu)", res[0], res[1]);
    return GPW_EUNSAT;
  }
  return GPW_OK;
}

extern "C" int gpw_r1cs_eval_dev(gpw_circuit* c, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                                 uint64_t* n_unsatisfied) {
  return gpw_r1cs_eval_on(c, c ? c->ctx : nullptr, wires_dev, a_dev, b_dev, c_dev, n_unsatisfied);
}

// variables.DeserializeProofWithPublicInputs + DeserializeVerifierOnlyCircuitData (variables/deserialize.go:114-156):
// JSON -> flat canonical input vector in circuit-input order. out: (n_public + n_secret) x 4 u64.
extern "C" int gpw_circuit_parse_inputs(const gpw_circuit* c, const char* proof_with_public_inputs_json,
                                        const char* verifier_only_circuit_data_json, uint64_t* out, size_t out_cap_u64) {
  if (!c || !proof_with_public_inputs_json || !verifier_only_circuit_data_json || !out) {
    set_error("parse_inputs: null argument");
    return GPW_EINVAL;
  }
  if (!c->is_verifier) {
    set_error("parse_inputs: not a verifier circuit");
    return GPW_EINVAL;
  }
  try {
    gadgets::InputValues iv = gadgets::ParseProofInputs(c->cd, proof_with_public_inputs_json, verifier_only_circuit_data_json);
    if (!c->baked.empty()) {
      // a bound circuit: the baked values must be the documents' (else this proof belongs to another inner circuit / is
      // another proof than the one compiled in) and are not inputs
      if (iv.sec.size() < c->baked.size() || !std::equal(c->baked.begin(), c->baked.end(), iv.sec.begin())) {
        set_error("parse_inputs: the documents do not match the values this circuit was compiled with (verifier-only data / proof)");
        return GPW_EINVAL;
      }
      iv.sec.erase(iv.sec.begin(), iv.sec.begin() + c->baked.size());
    }
    if (iv.pub.size() != c->api.NumPublic() || iv.sec.size() != c->api.NumSecret()) {
      set_error("parse_inputs: proof shape does not match the compiled circuit");
      return GPW_EINVAL;
    }
    if ((iv.pub.size() + iv.sec.size()) * 4 > out_cap_u64) {
      set_error("parse_inputs: output buffer too small");
      return GPW_EINVAL;
    }
    memcpy(out, iv.pub.data(), iv.pub.size() * 32);
    memcpy(out + iv.pub.size() * 4, iv.sec.data(), iv.sec.size() * 32);
  } catch (const std::exception& e) {
    set_error("parse_inputs: %s", e.what());
    return GPW_EINVAL;
  }
  return GPW_OK;
}

// Wires that occur in some L row (side 0) or R row (side 1) of the R1CS, ascending: the supports of the A / B
// proving-key bases (gnark filters the others out via pk.InfinityA / pk.InfinityB).
extern "C" int gpw_circuit_supports(const gpw_circuit* c, int side, uint32_t* out, size_t cap, size_t* n) {
  if (!c || !out || !n || side < 0 || side > 1) return GPW_EINVAL;
  std::vector<uint8_t> seen(c->api.NumWires(), 0);
  const auto& cons = c->api.Constraints();
  const auto& off = c->api.LeOffsets();
  const auto& wi = c->api.LeWires();
  for (size_t k = 0; k < cons.size() / 3; k++) {
    uint32_t le = cons[3 * k + side];
    for (uint32_t t = off[le]; t < off[le + 1]; t++) seen[wi[t]] = 1;
  }
  size_t cnt = 0;
  for (uint32_t w = 0; w < seen.size(); w++)
    if (seen[w]) {
      if (cnt >= cap) return GPW_EINVAL;
      out[cnt++] = w;
    }
  *n = cnt;
  return GPW_OK;
}

// Output wires of every instruction of one opcode, in creation (= reference call) order.
extern "C" int gpw_circuit_hint_wires(const gpw_circuit* c, int op, uint32_t* out, size_t cap, size_t* n) {
  if (!c || !n) return GPW_EINVAL;
  size_t cnt = 0;
  // the builder's log of hint calls: hints fused into a macro instruction are no longer tape instructions of their own
  for (const auto& h : c->api.HintLog()) {
    if (h.first != op) continue;
    if (out) {
      if (cnt >= cap) return GPW_EINVAL;
      out[cnt] = h.second;
    }
    cnt++;
  }
  *n = cnt;
  return GPW_OK;
}

// Stand-alone gadget circuits, the shapes of the reference's own unit-test circuits:
//   "poseidon_gl"     12 public outputs, 12 secret inputs   (poseidon/goldilocks_test.go:15-35)
//   "poseidon_bn254"  4 public outputs, 4 secret inputs     (poseidon/bn254_test.go:14-29)
//   "qe_mul_div"      public (a*b, a/b), secret a, b        (goldilocks/quadratic_extension_test.go)
//   "range_check"     one secret input                      (goldilocks/base_test.go:16-24)
extern "C" int gpw_circuit_compile_gadget(gpw_ctx* ctx, const char* name, gpw_circuit** out) {
  if (!ctx || !name || !out) {
    set_error("circuit_compile_gadget: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  gpw_circuit* c = new gpw_circuit();
  c->ctx = ctx;
  fe::API* api = &c->api;
  std::string nm(name);
  try {
    gadgets::GlChip gl(api);
    std::vector<fe::Variable> pub, sec;
    if (nm.rfind("gate:", 0) == 0) {  // one gate of plonk/gates as its own circuit (finalised by DefineGateCircuit)
      gadgets::DefineGateCircuit(api, nm.substr(5));
      int rc_gate = finish_compile(c);
      if (rc_gate != GPW_OK) {
        gpw_circuit_free(c);
        return rc_gate;
      }
      *out = c;
      return GPW_OK;
    }
    if (nm == "poseidon_gl") {
      for (int i = 0; i < 12; i++) pub.push_back(api->PublicInput());
      for (int i = 0; i < 12; i++) sec.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::PoseidonGlChip p(api);
      gadgets::GlState st;
      for (int i = 0; i < 12; i++) st[i] = sec[i];
      st = p.Poseidon(st);
      for (int i = 0; i < 12; i++) api->AssertIsEqual(st[i], pub[i]);
    } else if (nm == "poseidon_bn254") {
      for (int i = 0; i < 4; i++) pub.push_back(api->PublicInput());
      for (int i = 0; i < 4; i++) sec.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::PoseidonBn254Chip p(api);
      auto st = p.Poseidon({sec[0], sec[1], sec[2], sec[3]});
      for (int i = 0; i < 4; i++) api->AssertIsEqual(st[i], pub[i]);
    } else if (nm == "qe_mul_div") {
      for (int i = 0; i < 4; i++) pub.push_back(api->PublicInput());
      for (int i = 0; i < 4; i++) sec.push_back(api->SecretInput());
      api->EndInputs();
      gadgets::QE a = {sec[0], sec[1]}, b = {sec[2], sec[3]};
      gadgets::QE m = gl.MulExtension(a, b);
      auto d = gl.DivExtension(a, b);
      api->AssertIsEqual(m[0], pub[0]);
      api->AssertIsEqual(m[1], pub[1]);
      api->AssertIsEqual(d.first[0], pub[2]);
      api->AssertIsEqual(d.first[1], pub[3]);
    } else if (nm == "range_check") {
      fe::Variable x = api->SecretInput();
      api->EndInputs();
      gl.RangeCheck(x);
    } else {
      throw std::runtime_error("unknown gadget circuit '" + nm + "'");
    }
    api->Finalize();
  } catch (const std::exception& e) {
    set_error("circuit_compile_gadget: %s", e.what());
    delete c;
    return GPW_EINVAL;
  }
  int rc = finish_compile(c);
  if (rc != GPW_OK) {
    gpw_circuit_free(c);
    return rc;
  }
  *out = c;
  return GPW_OK;
}

// ---- compile cache (the r1cs.WriteTo the reference had to comment out, benchmark.go:94-99) -----------------------------
// File = "GPWF" | is_verifier | common_circuit_data.json | baked input values | fe::API blob (R1CS + scheduled tape).
extern "C" int gpw_circuit_save(const gpw_circuit* c, const char* path) {
  if (!c || !path) {
    set_error("circuit_save: null argument");
    return GPW_EINVAL;
  }
  std::ofstream os(path, std::ios::binary | std::ios::trunc);
  if (!os) {
    set_error("circuit_save: cannot open %s", path);
    return GPW_EINVAL;
  }
  const uint32_t magic = 0x46575047u, isv = c->is_verifier ? 1u : 0u;
  const uint64_t jl = c->common_json.size(), nb = c->baked.size();
  os.write((const char*)&magic, 4);
  os.write((const char*)&isv, 4);
  os.write((const char*)&jl, 8);
  os.write(c->common_json.data(), (std::streamsize)jl);
  os.write((const char*)&nb, 8);
  if (nb) os.write((const char*)c->baked.data(), (std::streamsize)(nb * 32));
  c->api.Serialize(os);
  os.flush();
  if (!os) {
    set_error("circuit_save: write to %s failed", path);
    return GPW_EINVAL;
  }
  return GPW_OK;
}

extern "C" int gpw_circuit_load(gpw_ctx* ctx, const char* path, gpw_circuit** out) {
  if (!ctx || !path || !out) {
    set_error("circuit_load: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  std::ifstream is(path, std::ios::binary);
  if (!is) {
    set_error("circuit_load: cannot open %s", path);
    return GPW_EINVAL;
  }
  gpw_circuit* c = new gpw_circuit();
  c->ctx = ctx;
  try {
    uint32_t magic = 0, isv = 0;
    uint64_t jl = 0, nb = 0;
    is.read((char*)&magic, 4);
    is.read((char*)&isv, 4);
    is.read((char*)&jl, 8);
    if (!is || magic != 0x46575047u || jl > (1u << 28)) throw std::runtime_error("not a gpw circuit file");
    c->common_json.resize(jl);
    is.read(&c->common_json[0], (std::streamsize)jl);
    is.read((char*)&nb, 8);
    if (!is || nb > (1u << 24)) throw std::runtime_error("truncated header");
    c->baked.resize(nb);
    if (nb) is.read((char*)c->baked.data(), (std::streamsize)(nb * 32));
    c->is_verifier = isv != 0;
    if (c->is_verifier) c->cd = gadgets::ReadCommonCircuitData(c->common_json);
    c->api.Deserialize(is);
    if (!c->api.Scheduled()) throw std::runtime_error("the cached circuit was saved before scheduling");
  } catch (const std::exception& e) {
    set_error("circuit_load: %s", e.what());
    delete c;
    return GPW_EINVAL;
  }
  int rc = finish_compile(c);  // device upload; the tape keeps the schedule it was saved with
  if (rc != GPW_OK) {
    gpw_circuit_free(c);
    return rc;
  }
  *out = c;
  return GPW_OK;
}

// Internal (not part of the C ABI): the host-side compiled circuit, for the trusted setup (csrc/setup.cu) and the
// circuit cache.
const gpw::fe::API* gpw_circuit_api_internal(const gpw_circuit* c) { return c ? &c->api : nullptr; }
