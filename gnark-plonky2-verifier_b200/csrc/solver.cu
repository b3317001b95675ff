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
#include "gl.cuh"
#include "host/frontend.h"
#include "host/gadgets.h"

namespace gpw {

using fe::NO_LE;

struct DInstr {
  uint32_t op_nout;  // op | nout << 8
  uint32_t out;
  uint32_t le[3];
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
};

enum SegKind { SEG_NARROW, SEG_WIDE, SEG_COUNT, SEG_COMMIT };
struct Segment {
  SegKind kind;
  uint32_t lo, hi;  // level range [lo, hi)
};

constexpr uint32_t WIDE_THRESHOLD = 8192;
constexpr int NARROW_THREADS = 128;

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

__device__ __forceinline__ Fr eval_le(const DevCircuit& c, const Fr* W, uint32_t le) {
  Fr acc = Fr::zero();
  const uint32_t e = c.le_off[le + 1];
  for (uint32_t k = c.le_off[le]; k < e; k++) {
    const uint32_t cid = c.le_coeff[k];
    const Fr v = ld_w(W + c.le_wire[k]);
    if (cid == 0) acc = add(acc, v);
    else if (cid == 1) acc = sub(acc, v);
    else acc = add(acc, mul(ld_w(c.coeffs + cid), v));
  }
  return acc;
}

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

// error codes written to err[proof] (first error wins)
constexpr int ERR_MULADD = 1, ERR_GLINV = 2, ERR_SPLIT = 3, ERR_BITS = 4, ERR_DECOMP = 5, ERR_DIV0 = 6;

__device__ void exec_instr(const DevCircuit& c, Fr* W, const DInstr& in, int* err, uint32_t* hist) {
  const uint32_t op = in.op_nout & 0xffu, nout = in.op_nout >> 8;
  switch (op) {
    case fe::OP_MUL: {
      Fr r = mul(eval_le(c, W, in.le[0]), eval_le(c, W, in.le[1]));
      if (in.le[2] != NO_LE) r = add(r, eval_le(c, W, in.le[2]));
      st_w(W + in.out, r);
      break;
    }
    case fe::OP_HINT_MULADD: {
      uint64_t a[4], b[4], d[4];
      to_u64x4(eval_le(c, W, in.le[0]), a);
      to_u64x4(eval_le(c, W, in.le[1]), b);
      to_u64x4(eval_le(c, W, in.le[2]), d);
      if ((a[1] | a[2] | a[3] | b[1] | b[2] | b[3] | d[1] | d[2] | d[3]) || a[0] >= gl::P || b[0] >= gl::P || d[0] >= gl::P) {
        atomicCAS(err, 0, ERR_MULADD);  // goldilocks/base.go:228-232 panics
        return;
      }
      uint64_t q, r;
      gl::mul_add_hint(a[0], b[0], d[0], q, r);
      st_w(W + in.out, from_u64(q));
      st_w(W + in.out + 1, from_u64(r));
      break;
    }
    case fe::OP_HINT_REDUCE: {
      uint64_t x[4], q[4], r;
      to_u64x4(eval_le(c, W, in.le[0]), x);
      gl::reduce_hint(x, q, r);
      st_w(W + in.out, from_u64x4(q));
      st_w(W + in.out + 1, from_u64(r));
      break;
    }
    case fe::OP_HINT_GLINV: {
      uint64_t x[4];
      to_u64x4(eval_le(c, W, in.le[0]), x);
      if ((x[1] | x[2] | x[3]) || x[0] >= gl::P) {
        atomicCAS(err, 0, ERR_GLINV);
        return;
      }
      st_w(W + in.out, from_u64(gl::inverse(x[0])));
      break;
    }
    case fe::OP_HINT_SPLIT: {
      uint64_t x[4];
      to_u64x4(eval_le(c, W, in.le[0]), x);
      if ((x[1] | x[2] | x[3]) || x[0] >= gl::P) {
        atomicCAS(err, 0, ERR_SPLIT);  // goldilocks/base.go:347-349 returns an error
        return;
      }
      st_w(W + in.out, from_u64(x[0] >> 32));
      st_w(W + in.out + 1, from_u64(x[0] & 0xffffffffull));
      break;
    }
    case fe::OP_INVZERO: st_w(W + in.out, inv(eval_le(c, W, in.le[0]))); break;
    case fe::OP_BITS: {
      uint64_t x[4];
      to_u64x4(eval_le(c, W, in.le[0]), x);
      const Fr one = Fr::one(), zero = Fr::zero();
      for (uint32_t i = 0; i < nout; i++) st_w(W + in.out + i, ((x[i >> 6] >> (i & 63)) & 1ull) ? one : zero);
      for (uint32_t i = nout; i < 256; i++)
        if ((x[i >> 6] >> (i & 63)) & 1ull) atomicCAS(err, 0, ERR_BITS);
      break;
    }
    case fe::OP_DIV: {
      Fr d = eval_le(c, W, in.le[1]);
      if (d.is_zero()) atomicCAS(err, 0, ERR_DIV0);
      st_w(W + in.out, mul(eval_le(c, W, in.le[0]), inv(d)));
      break;
    }
    case fe::OP_DECOMP: {
      uint64_t x[4];
      to_u64x4(eval_le(c, W, in.le[0]), x);
      for (uint32_t i = 0; i < nout; i++) {
        uint32_t v = (uint32_t)((x[(16 * i) >> 6] >> ((16 * i) & 63)) & 0xffffull);
        st_w(W + in.out + i, from_u64(v));
        atomicAdd(&hist[v], 1u);
      }
      for (uint32_t i = 16 * nout; i < 256; i += 16)
        if ((x[i >> 6] >> (i & 63)) & 0xffffull) atomicCAS(err, 0, ERR_DECOMP);  // value exceeds its range: unsatisfiable
      break;
    }
    default: break;
  }
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

// grid.y = proof
__global__ void __launch_bounds__(128)
    k_tape_wide(DevCircuit c, Fr* __restrict__ wires, size_t wire_stride, int* __restrict__ err, uint32_t* __restrict__ hist,
                uint32_t lvl) {
  Fr* W = wires + (size_t)blockIdx.y * wire_stride;
  const uint32_t s = c.level_off[lvl], t = c.level_off[lvl + 1];
  for (uint32_t i = s + blockIdx.x * blockDim.x + threadIdx.x; i < t; i += gridDim.x * blockDim.x)
    exec_instr(c, W, c.instr[i], err + blockIdx.y, hist + (size_t)blockIdx.y * 65536);
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

// a = L.w, b = R.w, c = O.w for every constraint (zero-padded to the FFT domain by the caller's memset);
// counts rows with a*b != c.
__global__ void __launch_bounds__(128)
    k_r1cs_eval(DevCircuit c, const Fr* __restrict__ W, Fr* __restrict__ a, Fr* __restrict__ b, Fr* __restrict__ cc,
                unsigned long long* __restrict__ n_bad, unsigned long long* __restrict__ first_bad) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= c.n_cons) return;
  Fr l = eval_le(c, W, c.cons[3 * k]), r = eval_le(c, W, c.cons[3 * k + 1]), o = eval_le(c, W, c.cons[3 * k + 2]);
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
  uint32_t n_inputs = 0;
  float solve_ms = 0;
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
  api.ScheduleALAP();
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
    di.push_back({(uint32_t)in.op | (in.nout << 8), in.out, {in.le[0], in.le[1], in.le[2]}});
    level_off[in.level + 1]++;
  }
  for (uint32_t l = 0; l < L; l++) level_off[l + 1] += level_off[l];
  // execution plan
  c->plan.clear();
  uint32_t run_lo = 0;
  auto flush = [&](uint32_t upto) {
    if (upto > run_lo) c->plan.push_back({SEG_NARROW, run_lo, upto});
  };
  for (uint32_t l = 0; l < L; l++) {
    const uint32_t cnt = level_off[l + 1] - level_off[l];
    const bool special = (l == count_level || l == commit_level);
    if (cnt >= WIDE_THRESHOLD || special) {
      flush(l);
      if (cnt) c->plan.push_back({SEG_WIDE, l, l + 1});
      if (l == count_level) c->plan.push_back({SEG_COUNT, l, l + 1});
      if (l == commit_level) c->plan.push_back({SEG_COMMIT, l, l + 1});
      run_lo = l + 1;
    }
  }
  flush(L);
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
  c->n_inputs = api.NumPublic() + api.NumSecret();
  return GPW_OK;
}

extern "C" void gpw_circuit_free(gpw_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  for (void* p : c->dev_allocs) cudaFree(p);
  delete c;
}

extern "C" int gpw_circuit_compile_verifier(gpw_ctx* ctx, const char* common_circuit_data_json, gpw_circuit** out) {
  if (!ctx || !common_circuit_data_json || !out) {
    set_error("circuit_compile: null argument");
    return GPW_EINVAL;
  }
  GPW_CUDA(cudaSetDevice(ctx->device));
  gpw_circuit* c = new gpw_circuit();
  c->ctx = ctx;
  try {
    c->cd = gadgets::ReadCommonCircuitData(common_circuit_data_json);
    gadgets::DefineVerifierCircuit(&c->api, c->cd);
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
static int run_segments(gpw_circuit* c, Fr* wires, size_t stride, int n_proofs, int* err, uint32_t* hist, size_t seg_lo,
                        size_t seg_hi) {
  gpw_ctx* ctx = c->ctx;
  cudaStream_t st = ctx->stream;
  for (size_t si = seg_lo; si < seg_hi; si++) {
    const Segment& s = c->plan[si];
    if (s.kind == SEG_NARROW) {
      k_tape_narrow<<<n_proofs, NARROW_THREADS, 0, st>>>(c->dc, wires, stride, err, hist, s.lo, s.hi);
      GPW_CHECK_LAUNCH();
      ctx->launches++;
    } else if (s.kind == SEG_WIDE) {
      uint32_t cnt = 0;
      // level width is known on the host from the plan construction; recompute cheaply from the API tape size bound
      cnt = 0;
      (void)cnt;
      dim3 grid(ctx->sm_count * 8, n_proofs);
      k_tape_wide<<<grid, 128, 0, st>>>(c->dc, wires, stride, err, hist, s.lo);
      GPW_CHECK_LAUNCH();
      ctx->launches++;
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

static int check_err(gpw_circuit* c, int* err_dev, int n_proofs) {
  std::vector<int> e(n_proofs);
  GPW_CUDA(cudaMemcpyAsync(e.data(), err_dev, n_proofs * sizeof(int), cudaMemcpyDeviceToHost, c->ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(c->ctx->stream));
  for (int i = 0; i < n_proofs; i++)
    if (e[i]) {
      set_error("witness solve failed for proof %d: %s", i, err_name(e[i]));
      return GPW_EHINT;
    }
  return GPW_OK;
}

// Phase 1: everything up to (not including) the commitment challenge. inputs_dev: n_proofs x n_inputs x 4 u64
// canonical (public then secret). wires_dev: n_proofs x wire_stride Fr.
extern "C" int gpw_witness_solve_phase1_dev(gpw_circuit* c, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride) {
  if (!c || !inputs_dev || !wires_dev || n_proofs < 1 || wire_stride < c->dc.n_wires) {
    set_error("witness_solve: bad argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = c->ctx;
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
  GPW_TRY(run_segments(c, (Fr*)wires_dev, wire_stride, n_proofs, err, hist, 0, commit_segment(c)));
  return check_err(c, err, n_proofs);
}

// Phase 2: sets the commitment challenge (one canonical Fr per proof) and runs the rest of the tape.
extern "C" int gpw_witness_solve_phase2_dev(gpw_circuit* c, const uint64_t* challenges_canonical, int n_proofs, uint64_t wires_dev,
                                            size_t wire_stride) {
  if (!c || !wires_dev || n_proofs < 1) {
    set_error("witness_solve: bad argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = c->ctx;
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
  for (int p = 0; p < n_proofs; p++) {
    Fr x = fe::fr_from_limbs(challenges_canonical + 4 * p);
    GPW_CUDA(cudaMemcpyAsync((Fr*)wires_dev + (size_t)p * wire_stride + c->dc.commit_wire, &x, sizeof(Fr), cudaMemcpyHostToDevice,
                             ctx->stream));
    GPW_CUDA(cudaStreamSynchronize(ctx->stream));  // x lives on the host stack
  }
  GPW_TRY(run_segments(c, (Fr*)wires_dev, wire_stride, n_proofs, err, hist, cs + 1, c->plan.size()));
  return check_err(c, err, n_proofs);
}

// a, b, c evaluation vectors of one proof (may be 0 to only check). Returns GPW_EUNSAT if a constraint fails.
extern "C" int gpw_r1cs_eval_dev(gpw_circuit* c, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                                 uint64_t* n_unsatisfied) {
  if (!c || !wires_dev) {
    set_error("r1cs_eval: null argument");
    return GPW_EINVAL;
  }
  gpw_ctx* ctx = c->ctx;
  GPW_CUDA(cudaSetDevice(ctx->device));
  unsigned long long* bad;
  GPW_TRY(ctx->get_scratch("solve.bad", 16, (void**)&bad));
  unsigned long long init[2] = {0, ~0ull};
  GPW_CUDA(cudaMemcpyAsync(bad, init, 16, cudaMemcpyHostToDevice, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  k_r1cs_eval<<<div_up(c->dc.n_cons, 128), 128, 0, ctx->stream>>>(c->dc, (const Fr*)wires_dev, (Fr*)a_dev, (Fr*)b_dev, (Fr*)c_dev, bad,
                                                                  bad + 1);
  GPW_CHECK_LAUNCH();
  ctx->launches++;
  unsigned long long res[2];
  GPW_CUDA(cudaMemcpyAsync(res, bad, 16, cudaMemcpyDeviceToHost, ctx->stream));
  GPW_CUDA(cudaStreamSynchronize(ctx->stream));
  if (n_unsatisfied) *n_unsatisfied = res[0];
  if (res[0]) {
    set_error("%llu constraints unsatisfied (first: #%llu)", res[0], res[1]);
    return GPW_EUNSAT;
  }
  return GPW_OK;
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
  for (const auto& in : c->api.Tape()) {
    if (in.op != op) continue;
    for (uint32_t i = 0; i < in.nout; i++) {
      if (out) {
        if (cnt >= cap) return GPW_EINVAL;
        out[cnt] = in.out + i;
      }
      cnt++;
    }
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
