/* libgpw - B200-native hot path of the Plonky2 -> gnark (Groth16 over BN254) wrap prover.
 *
 * C ABI (extern "C", plain pointers and sizes). This is the boundary the reference's Go code would
 * bind through cgo (INTEGRATION.md shows the stubs). Each entry point cites the reference interface
 * it replaces. Conventions:
 *   - every function returns 0 on success, a negative GPW_E* code on failure; gpw_last_error() gives
 *     the text (thread-local). No exceptions cross the boundary.
 *   - field elements are 4 x uint64 little-endian limbs. "mont" = Montgomery form with R = 2^256,
 *     i.e. the in-memory form of gnark-crypto's fr.Element / fp.Element; "canonical" =
 *     the plain integer.
 *   - G1 affine = {X, Y} fp (64 B), G2 affine = {X.A0, X.A1, Y.A0, Y.A1} fp (128 B), Montgomery form,
 *     (0,0) = infinity - gnark-crypto's G1Affine / G2Affine layout.
 *   - functions without a suffix take HOST buffers (copies are inside the call); functions with the
 *     _dev suffix take DEVICE addresses (as returned by cudaMalloc / torch's data_ptr()) and enqueue
 *     on the context's stream without synchronising unless stated.
 *   - a gpw_ctx is bound to one GPU; one ctx per GPU / per host thread. There is no CPU fallback:
 *     every compute entry point fails with GPW_ENODEV when no CUDA device is usable.
 */
#ifndef GPW_H
#define GPW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPW_OK 0
#define GPW_EINVAL (-1)
#define GPW_ENODEV (-2)
#define GPW_ECUDA (-3)
#define GPW_ENOMEM (-4)
#define GPW_EHINT (-5)   /* a hint rejected its input (reference: panic / error return) */
#define GPW_EUNSAT (-6)  /* witness does not satisfy the constraint system */
#define GPW_ENCCL (-7)

typedef struct gpw_ctx gpw_ctx;

/* ---- library / context -------------------------------------------------------------------- */
int gpw_version(void);
const char* gpw_last_error(void);
int gpw_device_count(void);
int gpw_ctx_create(int device, gpw_ctx** out);
void gpw_ctx_destroy(gpw_ctx* ctx);
/* Use an externally owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream. */
int gpw_ctx_set_stream(gpw_ctx* ctx, void* cuda_stream);
int gpw_ctx_sync(gpw_ctx* ctx);
/* Tunables. "msm_affine_rounds" (0..8, default 0): rounds of batch-affine pairwise bucket reduction - gnark-crypto's
 * batch-affine MultiExp restated for the GPU (csrc/msm_affine.cuh) - run before the extended-Jacobian accumulation.
 * Same results bit for bit; measured slower than XYZZ alone on B200, hence off (profiles/r02_batch_affine.md).
 * "msm_overlap" (-1 automatic = default, 0 off, 1 on), read from the KEY's context by the wrap entry points: deferred MSMs -
 * the latency-bound tail of one MSM of a proof runs on a high-priority stream beside the bucket accumulation of the next and
 * the host folds all window sums at the end. Automatic = on: ~10 ms shorter lone proofs, 13.2 -> 13.6 proofs/s for a stream
 * (the solve runs on the high-priority stream as well). Same proofs bit for bit either way.                              */
int gpw_ctx_set_option(gpw_ctx* ctx, const char* key, int64_t value);
/* Number of gpw kernels launched through this ctx since creation (bench.py's gpu_launches). */
uint64_t gpw_ctx_launch_count(const gpw_ctx* ctx);

/* ---- host-side arithmetic (no GPU): setup glue and the unit tests of the shared host/device code */
/* field: 0 = Fr, 1 = Fp. impl: 0 = even/odd IMAD.WIDE schedule (host emulation), 1 = plain CIOS.  */
int gpw_host_ff_mul(int field, int impl, const uint64_t* a_mont, const uint64_t* b_mont, uint64_t* out_mont, size_t n);
/* a b - c d with one Montgomery reduction for both products (the form the bucket accumulation uses for Y3) */
int gpw_host_ff_mul_sub2(int field, const uint64_t* a_mont, const uint64_t* b_mont, const uint64_t* c_mont, const uint64_t* d_mont,
                         uint64_t* out_mont, size_t n);
int gpw_host_ff_to_mont(int field, const uint64_t* a, uint64_t* out, size_t n);
int gpw_host_ff_from_mont(int field, const uint64_t* a, uint64_t* out, size_t n);
int gpw_host_ff_inv(int field, const uint64_t* a_mont, uint64_t* out_mont, size_t n);
/* same result through the binary extended Euclid the device uses for its batched inversions (csrc/ff.cuh inv_euclid) */
int gpw_host_ff_inv_euclid(int field, const uint64_t* a_mont, uint64_t* out_mont, size_t n);
/* group: 1 = G1, 2 = G2. out = k * P with k a canonical 256-bit scalar; affine in/out, mont. */
int gpw_host_ec_scalar_mul(int group, const uint64_t* point_affine, const uint64_t* scalar_canonical, uint64_t* out_affine);
int gpw_host_ec_add(int group, const uint64_t* p_affine, const uint64_t* q_affine, uint64_t* out_affine);
int gpw_host_ec_is_on_curve(int group, const uint64_t* p_affine);
/* n points [k_i]G for k_i = k0 + i (fixed generator); used to make synthetic bases cheaply. */
int gpw_host_ec_generator_multiples(int group, uint64_t k0, size_t n, uint64_t* out_affine);

/* ---- GPU self-test of the field kernels: runs the PTX multiply and the portable multiply on the
 * device on n seeded pairs, compares both with the host result. Returns 0 if all agree. */
int gpw_selftest_ff(gpw_ctx* ctx, size_t n, uint64_t seed);

/* ---- K9: Pippenger MSM over BN254 G1 / G2 ---------------------------------------------------
 * Replaces gnark-crypto G1Jac.MultiExp / G2Jac.MultiExp (ecc/bn254, un-vendored; reached from
 * groth16.Prove at benchmark.go:249 and plonk.Prove at benchmark.go:162).
 * scalars: n x 4 u64 (Montgomery if scalars_mont != 0, else canonical); points: n affine, mont.
 * out_affine: the sum as one affine point (8 u64 for G1, 16 for G2), (0,0) for infinity.
 * window_bits: Pippenger window c in [4,16], 0 = choose from n.                                 */
int gpw_msm_g1(gpw_ctx* ctx, const uint64_t* scalars, const uint64_t* points, size_t n, int scalars_mont,
               int window_bits, uint64_t* out_affine);
int gpw_msm_g2(gpw_ctx* ctx, const uint64_t* scalars, const uint64_t* points, size_t n, int scalars_mont,
               int window_bits, uint64_t* out_affine);
/* Device-resident variants. Synchronise the stream before returning (the window sums are folded on
 * the host). win_lo/win_hi select a sub-range of windows [win_lo, win_hi) for the multi-GPU window
 * split (pass 0, 0 for all): the returned point is then sum_{w in range} 2^(c w) W_w.            */
int gpw_msm_g1_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                   int window_bits, int win_lo, int win_hi, uint64_t* out_affine);
/* Fixed-base MSM for bases known at setup (a proving key): gpw_msm_g1_fixed_table fills table_dev (n_windows * n affine
 * points, entry w n + i = 2^(window_bits w) P_i) with n_windows = ceil(255 / window_bits), window_bits <= 24;
 * gpw_msm_g1_fixed_dev then puts every digit of every scalar into ONE bucket set: ceil(255 / window_bits) bucket
 * additions per full-width scalar (12 at 22 bits instead of 16 at 16 bits) and one bucket reduction.                  */
int gpw_msm_g1_fixed_table(gpw_ctx* ctx, uint64_t points_dev, size_t n, int window_bits, int n_windows, uint64_t table_dev);
int gpw_msm_g1_fixed_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t table_dev, size_t n, int scalars_mont, int window_bits,
                         int n_windows, uint64_t* out_affine);
int gpw_msm_g2_dev(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                   int window_bits, int win_lo, int win_hi, uint64_t* out_affine);
/* ---- one MSM over several GPUs (BASELINE configs[4]; gnark-crypto's MultiExp splits its windows over CPU cores, this
 * splits them - or the points - over B200s). One process per GPU; rank 0 gets an id from gpw_comm_unique_id and hands the
 * 128 bytes to the other ranks through the host program's own channel; every rank then calls gpw_comm_init (collective,
 * ncclCommInitRank on the context's device). libnccl.so.2 is loaded lazily; GPW_ENCCL if it is missing or a call fails.  */
int gpw_comm_unique_id(uint8_t* out128);
int gpw_comm_init(gpw_ctx* ctx, int nranks, int rank, const uint8_t* id128);
int gpw_comm_destroy(gpw_ctx* ctx);
int gpw_comm_info(const gpw_ctx* ctx, int* info3); /* {ranks (0 = no communicator), this rank, NCCL version code} */
/* Collective over the context's communicator. split: 1 = every rank holds all n scalars / bases and computes its range of
 * Pippenger windows; 2 = rank r computes the points [n r / N, n (r + 1) / N) (it only touches that slice); 0 = choose. One
 * all-gather of one affine point per rank on the context's stream, the N points are added on the device; every rank
 * receives the full result, bit-identical to gpw_msm_g{1,2}_dev on one GPU.                                              */
int gpw_msm_g1_sharded(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                       int split, uint64_t* out_affine);
int gpw_msm_g2_sharded(gpw_ctx* ctx, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont, int window_bits,
                       int split, uint64_t* out_affine);
/* The share rank `rank` of `nranks` contributes under split 1 / 2, computed on this context without a communicator (group:
 * 1 = G1, 2 = G2): the shares of all ranks add up to the whole MSM.                                                      */
int gpw_msm_sharded_partial(gpw_ctx* ctx, int group, uint64_t scalars_dev, uint64_t points_dev, size_t n, int scalars_mont,
                            int window_bits, int split, int rank, int nranks, uint64_t* out_affine);
/* Time (ms, CUDA events on the ctx stream) spent in the bucket-accumulation kernel of the most
 * recent MSM on this ctx, and the number of non-zero digits it processed. */
int gpw_msm_last_stats(gpw_ctx* ctx, float* accumulate_ms, float* total_ms, uint64_t* nonzero_digits);
/* Cumulative per-group (1 = G1, 2 = G2) statistics since the last reset: {accumulate ms, whole-MSM ms, points,
 * non-zero digits, calls}; the figures bench.py derives roofline.achieved from. */
int gpw_msm_cumulative_stats(gpw_ctx* ctx, int group, int reset, double* out5);

/* ---- K8: radix-2 NTT over BN254 Fr ------------------------------------------------------------
 * Replaces gnark-crypto fr/fft Domain.FFT / FFTInverse (+ OnCoset) as used by groth16 computeH.
 * data: 2^logn Fr elements, Montgomery form, transformed in place.
 *   forward:  A_k = sum_j a_j w^(jk)            (coset: a_j <- a_j g^j first, g = 5)
 *   inverse:  a_j = N^-1 sum_k A_k w^(-jk)      (coset: a_j <- a_j g^-j afterwards)
 * in_bitrev / out_bitrev select bit-reversed index order on either side (DIF: nat->bitrev,
 * DIT: bitrev->nat; nat->nat adds one permutation pass).                                       */
int gpw_ntt_fr(gpw_ctx* ctx, uint64_t* data, int logn, int inverse, int coset, int in_bitrev, int out_bitrev);
int gpw_ntt_fr_dev(gpw_ctx* ctx, uint64_t data_dev, int logn, int inverse, int coset, int in_bitrev, int out_bitrev);
/* `to` (another context of the same device) borrows `from`'s twiddle / coset tables for 2^logn; `from` must outlive it. */
int gpw_ntt_share_tables(gpw_ctx* from, gpw_ctx* to, int logn);

/* ---- Groth16 prover glue (gnark backend/groth16 bn254 Prove, benchmark.go:249) -----------
 * computeH pointwise step on the coset: a[i] = (a[i] * b[i] - c[i]) * k, all Fr Montgomery, device.   */
int gpw_fr_h_pointwise_dev(gpw_ctx* ctx, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, size_t n, const uint64_t* k_mont);
/* In-place canonical <-> Montgomery conversion of n Fr elements on the device (to_mont != 0: into Montgomery). */
int gpw_fr_convert_dev(gpw_ctx* ctx, uint64_t a_dev, size_t n, int to_mont);
/* out[i] = [k0 + i] G (affine, mont) for the fixed generator of G1 (group 1) or G2 (group 2), written to
 * device memory: synthetic proving-key bases for benchmarks / tests (what gnark's DummySetup provides). */
int gpw_ec_generator_multiples_dev(gpw_ctx* ctx, int group, uint64_t k0, size_t n, uint64_t out_dev);

/* Proving key handle (device-resident bases). gnark: groth16.ProvingKey (benchmark.go:214-217).           */
typedef struct gpw_pk gpw_pk;
/* Synthetic key with the shapes / cost of a real one and KNOWN discrete logs (see csrc/groth16.cu): the
 * analogue of groth16.DummySetup (benchmark.go:214). m wires (wire 0 = constant one), the first n_pub are
 * public, FFT domain 2^logN.                                                                            */
int gpw_groth16_pk_synthetic(gpw_ctx* ctx, size_t m, size_t n_pub, int logN, uint64_t seed, gpw_pk** out);
void gpw_groth16_pk_free(gpw_pk* pk);
int gpw_groth16_pk_info(const gpw_pk* pk, uint64_t* m, uint64_t* n_pub, int* logN);
/* computeH: a <- coefficients of (A.B - C)/Z_H, natural order; a, b, c are N = 2^logN evaluations on H.  */
int gpw_groth16_compute_h_dev(gpw_ctx* ctx, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, int logN);
/* groth16.Prove after the solve (benchmark.go:249): computeH + the five MSMs + assembly. r, s canonical.
 * out = Ar (8 u64) | Bs (16 u64) | Krs (8 u64), affine Montgomery coordinates.                          */
int gpw_groth16_prove_dev(gpw_pk* pk, uint64_t w_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                          const uint64_t* r_canonical, const uint64_t* s_canonical, uint64_t* out_proof);
/* ms spent in computeH and in each of the MSMs {A, B1, B2, K, Z} of the last prove (CUDA events).        */
int gpw_groth16_last_stats(const gpw_pk* pk, float* h_ms, float* msm_ms5);

/* ---- Circuit compile + witness synthesis (K1-K7 as one levelled tape) ---------------------------------------
 * gpw_circuit = the compiled verifier circuit (R1CS + solver tape on the device). Replaces
 * frontend.Compile(ecc.BN254.ScalarField(), r1cs.NewBuilder, &circuit) at benchmark.go:55 for
 * verifier.ExampleVerifierCircuit (verifier/util.go:10-24); the proof and the verifier-only data are runtime
 * (secret) inputs as in the reference's own test circuits (fri/fri_test.go:17-21), PublicInputs are public.  */
typedef struct gpw_circuit gpw_circuit;
int gpw_circuit_compile_verifier(gpw_ctx* ctx, const char* common_circuit_data_json, gpw_circuit** out);
/* The reference's own forms of ExampleVerifierCircuit (verifier/util.go:10-24, `gnark:"-"` = compile-time constant):
 *   verifier_only_json != NULL: constants_sigmas_cap + circuit_digest are CONSTANTS of the circuit - the statement proven is
 *     "a proof of THIS inner circuit verifies for these public inputs" (the runtime-input form above proves it for whatever
 *     verifier data the prover supplies, so a key made from it must only be used where that is intended);
 *   proof_json != NULL too: the proof is baked in as well - literally what benchmark.go:33-55 compiles.
 * gpw_circuit_parse_inputs on such a circuit checks the documents against the baked values and omits them.          */
int gpw_circuit_compile_verifier_bound(gpw_ctx* ctx, const char* common_circuit_data_json, const char* verifier_only_json,
                                       const char* proof_json, gpw_circuit** out);
/* Stand-alone gadget circuits shaped like the reference's unit-test circuits: "poseidon_gl", "poseidon_bn254",
 * "qe_mul_div", "range_check" (poseidon/goldilocks_test.go, poseidon/bn254_test.go, goldilocks/*_test.go).       */
int gpw_circuit_compile_gadget(gpw_ctx* ctx, const char* name, gpw_circuit** out);
/* also: "gate:<n_consts>:<n_wires>:<n_constraints>:<gate id>" - ONE gate of plonk/gates (gate id as in common_circuit_data.json)
 * as a circuit: secret inputs = constants, wires (2 limbs each), public-inputs hash (4); public inputs = the expected value of
 * every constraint (2 limbs each); satisfied iff Gate.EvalUnfiltered gives those values (plonk/gates/gates_test.go's check). */
void gpw_circuit_free(gpw_circuit* c);
/* Compile cache: the compiled circuit (R1CS + scheduled solver tape + input codec state) as one file, so that every rank and
 * every later process loads in a fraction of the compile time instead of re-running the gadget code. This is the
 * `r1cs.WriteTo(fR1CS)` the reference had to comment out for this circuit ("takes up too much memory", benchmark.go:94-99,
 * 204-209). The file belongs to this build of libgpw (layout-versioned); a foreign or damaged file is refused.          */
int gpw_circuit_save(const gpw_circuit* c, const char* path);
int gpw_circuit_load(gpw_ctx* ctx, const char* path, gpw_circuit** out);
/* info16: wires, public, secret, constraints, instructions, levels, limb_wires, limb_start, count_start, commit_wire,
 * narrow segments, wide segments, #MulAddHint, #ReduceHint, #InverseHint, #SplitLimbsHint                          */
int gpw_circuit_info(const gpw_circuit* c, uint64_t* info16);
/* types.ReadProofWithPublicInputs + variables.Deserialize* (types/deserialize.go:92, variables/deserialize.go:114-156):
 * the two JSON documents -> flat canonical input vector (public then secret, 4 u64 each).                          */
int gpw_circuit_parse_inputs(const gpw_circuit* c, const char* proof_with_public_inputs_json,
                             const char* verifier_only_circuit_data_json, uint64_t* out, size_t out_cap_u64);
/* The solve of groth16.Prove (benchmark.go:249): phase 1 = everything before the range-check commitment challenge,
 * phase 2 = the log-derivative argument after it. inputs_dev: n_proofs x n_inputs x 4 u64 canonical; wires_dev:
 * n_proofs x wire_stride Fr (Montgomery). challenges: one canonical Fr per proof (NULL if the circuit has none). */
int gpw_witness_solve_phase1_dev(gpw_circuit* c, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_witness_solve_phase2_dev(gpw_circuit* c, const uint64_t* challenges_canonical, int n_proofs, uint64_t wires_dev,
                                 size_t wire_stride);
/* a = L.w, b = R.w, c = O.w (device, >= n_constraints Fr each; pass 0 to only check). GPW_EUNSAT if a*b != c somewhere. */
int gpw_r1cs_eval_dev(gpw_circuit* c, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev, uint64_t* n_unsatisfied);
/* The same three calls on `lane`, any other context of the circuit's device (a context = a stream + its scratch
 * memory): the compiled circuit is read-only, so several proofs can be solved side by side, one per lane.          */
int gpw_witness_solve_phase1_on(gpw_circuit* c, gpw_ctx* lane, uint64_t inputs_dev, int n_proofs, uint64_t wires_dev, size_t wire_stride);
int gpw_witness_solve_phase2_on(gpw_circuit* c, gpw_ctx* lane, const uint64_t* challenges_canonical, int n_proofs, uint64_t wires_dev,
                                size_t wire_stride);
int gpw_r1cs_eval_on(gpw_circuit* c, gpw_ctx* lane, uint64_t wires_dev, uint64_t a_dev, uint64_t b_dev, uint64_t c_dev,
                     uint64_t* n_unsatisfied);
int gpw_circuit_supports(const gpw_circuit* c, int side, uint32_t* out, size_t cap, size_t* n);
/* output wires of every tape instruction of one opcode (1 MulAddHint, 2 ReduceHint, 3 InverseHint, 4 SplitLimbsHint),
 * in the order the reference's gadget code requests them.                                                       */
int gpw_circuit_hint_wires(const gpw_circuit* c, int op, uint32_t* out, size_t cap, size_t* n);

/* ---- the whole wrap: JSON inputs -> Groth16 proof (benchmark.go:240-249 on the GPU) --------------------------------
 * gpw_wrap_key_synthetic = groth16.DummySetup(r1cs) (benchmark.go:214) for a compiled circuit, with known discrete
 * logs (csrc/wrap.cu). gpw_wrap_prove: inputs on the HOST; out_proof (64 u64): Ar (8) | Bs (16) | Krs (8) |
 * commitment (8) | commitment PoK (8) | challenge (4, canonical) | n_unsatisfied (1).                             */
typedef struct gpw_wrap_key gpw_wrap_key;
int gpw_wrap_key_synthetic(gpw_ctx* ctx, gpw_circuit* circ, uint64_t seed, gpw_wrap_key** out);
void gpw_wrap_key_free(gpw_wrap_key* k);
int gpw_wrap_key_info(const gpw_wrap_key* k, uint64_t* info8);
/* groth16.Setup(r1cs) (benchmark.go:217): a REAL key for the compiled circuit - toxic waste tau, alpha, beta, gamma, delta
 * (+ the Pedersen key of the range-check commitment), QAP polynomials at tau from the R1CS, every base a multiple of the
 * generators (csrc/setup.cu). seed32: 32 bytes the toxic waste is derived from (reproducible keys for tests); NULL = fresh
 * OS entropy. Proofs made with such a key satisfy gnark's Groth16 verification equation (oracle/pairing.py checks it). */
int gpw_wrap_key_setup(gpw_ctx* ctx, gpw_circuit* circ, const uint8_t* seed32, gpw_wrap_key** out);
/* pk.WriteRawTo + vk.WriteRawTo (benchmark.go:224-232) in gnark's field order with uncompressed big-endian points, and the
 * inverse. The file describes THIS library's R1CS of the circuit (own wire numbering); shapes are checked on load.
 * vk_path may be NULL.                                                                                              */
int gpw_wrap_key_save(const gpw_wrap_key* k, const char* pk_path, const char* vk_path);
int gpw_wrap_key_load(gpw_ctx* ctx, gpw_circuit* circ, const char* pk_path, const char* vk_path, gpw_wrap_key** out);
/* vk.WriteRawTo into a caller buffer: alpha1 | beta1 | beta2 | gamma2 | delta1 | delta2 | u32 len(K) | K | u32 n_commitments |
 * per commitment u32 n + n x u64 | Pedersen vk (G | GRootSigmaNeg). out == NULL: only *len is set.                  */
int gpw_wrap_key_vk_write_raw(const gpw_wrap_key* k, uint8_t* out, size_t cap, size_t* len);
/* proof.WriteRawTo (benchmark.go:272-274): the 64-word proof of gpw_wrap_prove* as gnark's raw proof bytes
 * Ar (64) | Bs (128: X.A1 X.A0 Y.A1 Y.A0) | Krs (64) | u32 n_commitments | commitments | CommitmentPok (64), big-endian
 * canonical coordinates - benchmark.go:283-290 reads a, b, c from the first 256 bytes. out == NULL: only *len is set. */
int gpw_wrap_proof_write_raw(const gpw_wrap_key* k, const uint64_t* proof64, uint8_t* out, size_t cap, size_t* len);
uint64_t gpw_wrap_key_wires_dev(const gpw_wrap_key* k);
/* device address of the quotient coefficients h_0 .. h_{N-2} (Fr, Montgomery) left by the last gpw_wrap_prove (test aid) */
uint64_t gpw_wrap_key_h_dev(const gpw_wrap_key* k);
/* r_canonical / s_canonical: the prover's blinding scalars. NULL (the production path) = drawn from the OS CSPRNG inside
 * the library for every proof, as gnark's Prove does (crypto/rand); non-NULL = caller-supplied, for reproducible proofs in
 * tests. An unsatisfied constraint system never yields a proof: the call returns GPW_EUNSAT (count in slot 52) whatever
 * `check` is.                                                                                                        */
int gpw_wrap_prove(gpw_wrap_key* k, const uint64_t* inputs, const uint64_t* r_canonical, const uint64_t* s_canonical, int check,
                   uint64_t* out_proof);
/* same, with the parsed inputs already resident on the device (n_inputs x 4 u64 canonical) */
int gpw_wrap_prove_dev(gpw_wrap_key* k, uint64_t inputs_dev, const uint64_t* r_canonical, const uint64_t* s_canonical, int check,
                       uint64_t* out_proof);
/* A stream of n independent proofs with several of them in flight: one host thread + stream + scratch ("lane") per
 * in-flight proof, so that the sequential solve spine of one proof (one SM) and the host glue of another overlap the
 * MSMs / NTTs of the rest. inputs: n x n_inputs x 4 u64 (host or device); r, s: n x 4 u64; out: n x 64 u64. Blocking.
 * gpw_wrap_set_lanes: proofs in flight (default 6, env GPW_WRAP_LANES; 1 = strictly one after the other).           */
int gpw_wrap_prove_many(gpw_wrap_key* k, const uint64_t* inputs, int n, const uint64_t* r_canonical, const uint64_t* s_canonical,
                        int check, uint64_t* out_proofs);
int gpw_wrap_set_lanes(gpw_wrap_key* k, int n);
int gpw_wrap_last_stats(const gpw_wrap_key* k, float* ms6);
int gpw_hash_to_fr(const uint8_t* msg, size_t len, const char* dst, uint64_t* out_canonical);

/* ---- PLONK / KZG backend (BASELINE configs[3]; `-proof-system plonk`, benchmark.go:80-190) ---------------------------------
 * gpw_plonk_setup = test.NewKZGSRS + plonk.Setup (benchmark.go:105, 130) for a compiled circuit: the circuit is lowered to PLONK
 * gates (csrc/host/scs.cc: addition chains for the linear expressions, one multiplication gate per R1CS row, public-input rows,
 * one Qcp row per committed range-check wire), the permutation and selector polynomials are committed with a KZG SRS generated
 * from seed32 (NULL = OS entropy). gpw_plonk_prove = frontend.NewWitness + plonk.Prove (benchmark.go:162): witness on the GPU,
 * five rounds of PLONK with the BSB22 commitment column, proof = 10 G1 commitments (64 B raw big-endian each: a b c P2 Z t0 t1 t2
 * W_zeta W_zeta_w) + 18 evaluations (32 B big-endian each). The protocol is the published one; it differs from gnark's
 * implementation in what is opened and in the transcript labels (header of csrc/plonk.cu) and has no blinding (not
 * zero-knowledge). oracle/plonk_verify.py is its verifier. GPW_EUNSAT if the witness does not satisfy the system.                    */
typedef struct gpw_plonk_key gpw_plonk_key;
int gpw_plonk_setup(gpw_ctx* ctx, gpw_circuit* circ, const uint8_t* seed32, gpw_plonk_key** out);
void gpw_plonk_key_free(gpw_plonk_key* k);
/* info8: logN, gates, variables, public rows (ONE + public inputs + commitment challenge), Qcp rows, inputs, has_commit, chain levels */
int gpw_plonk_key_info(const gpw_plonk_key* k, uint64_t* info8);
/* vk: u32 logN | u32 n_public_rows | u32 has_commit | k1 | k2 | omega | [qL] [qR] [qM] [qO] [qC] [Qcp] [S1] [S2] [S3] | [tau]2 */
int gpw_plonk_vk_write(const gpw_plonk_key* k, uint8_t* out, size_t cap, size_t* len);
/* inputs: n_inputs x 4 u64 canonical, host (gpw_circuit_parse_inputs order). out_proof: 1216 bytes. */
int gpw_plonk_prove(gpw_plonk_key* k, const uint64_t* inputs, uint8_t* out_proof, size_t cap);
/* ms of the last prove: witness + P2, wire columns, grand product, quotient, evaluations, openings */
int gpw_plonk_last_stats(const gpw_plonk_key* k, float* ms6);

/* ---- K3: Poseidon over BN254 Fr (t = 4) ---------------------------------------------------------
 * Replaces poseidon.BN254Chip.Poseidon (poseidon/bn254.go:39-45). states: n x 4 Fr in / out.   */
int gpw_poseidon_bn254(gpw_ctx* ctx, const uint64_t* states_in, uint64_t* states_out, size_t n, int mont);
int gpw_poseidon_bn254_dev(gpw_ctx* ctx, uint64_t in_dev, uint64_t out_dev, size_t n, int mont);
/* Merkle path walk of fri.verifyMerkleProofToCapWithCapIndex (fri/fri.go:97-144): for each of n paths,
 * leaf digest (Fr) + depth siblings + index bits (LSB first, as a u64) -> root. canonical in/out. */
int gpw_merkle_paths_bn254(gpw_ctx* ctx, const uint64_t* leaf_digests, const uint64_t* siblings, const uint64_t* index_bits,
                           size_t n_paths, int depth, uint64_t* roots_out);
/* BN254Chip.HashOrNoop (poseidon/bn254.go:47-94): n leaves of leaf_len Goldilocks elements -> digest. */
int gpw_hash_or_noop_bn254(gpw_ctx* ctx, const uint64_t* leaves, size_t n, int leaf_len, uint64_t* digests_out);

/* ---- K1: the four Goldilocks solver hints (goldilocks/base.go:223,284,316,339) -------------------
 * Batched; bit-exact (q, r). mul_add: inputs a,b,c (n each, u64 < p) -> q,r (u64).
 * reduce: x as 4 x u64 canonical (n x 4) -> q (n x 4, only the low 192 bits can be non-zero), r.
 * Returns GPW_EHINT if any input violates the reference's precondition (first bad index in
 * gpw_last_error()).                                                                            */
int gpw_gl_mul_add_hint(gpw_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* c, size_t n, uint64_t* q, uint64_t* r);
int gpw_gl_reduce_hint(gpw_ctx* ctx, const uint64_t* x4, size_t n, uint64_t* q4, uint64_t* r);
int gpw_gl_inverse_hint(gpw_ctx* ctx, const uint64_t* x, size_t n, uint64_t* inv);
int gpw_gl_split_limbs_hint(gpw_ctx* ctx, const uint64_t* x, size_t n, uint64_t* hi, uint64_t* lo);

/* ---- K2: Poseidon over Goldilocks (width 12) ------------------------------------------------------
 * Replaces poseidon.GoldilocksChip.Poseidon (poseidon/goldilocks.go:30-37). states: n x 12 u64.
 * (plain permutation; the per-hint (q, r) witness trace is produced by the tape executor.)        */
int gpw_poseidon_gl(gpw_ctx* ctx, const uint64_t* states_in, uint64_t* states_out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* GPW_H */
