/*
 * zkcert_cuda.h — C ABI of libzkcert_cuda.so: the B200 (sm_100a) implementation of the halo2-axiom
 * `create_proof` hot path that zkCert/halo2-zkcert drives.
 *
 * The reference has no FFI for this path: its boundary is Rust generics inside the un-vendored
 * `halo2_proofs` 0.2.0 crate (axiom fork @4b42325, /root/reference/Cargo.lock:1320-1336) and
 * `halo2curves` 0.4.0 (@e185711, Cargo.lock:1359-1380).  Each entry point below names the Rust
 * operator it replaces and the reference call sites that reach it.  INTEGRATION.md
 * shows the `extern "C"` block a maintainer adds to the patched halo2_proofs.
 *
 * Conventions
 *   - zkc_fr / zkc_fq: 32 bytes, little-endian limbs of the Montgomery residue (R = 2^256): the
 *     exact memory of halo2curves `Fr` / `Fq`, so `&[Fr]` passes as `const zkc_fr*` unconverted.
 *   - zkc_g1_affine: (x, y), 64 bytes, identity = (0, 0) — halo2curves `G1Affine`.
 *   - zkc_g1: Jacobian (x, y, z), 96 bytes — halo2curves `G1`.  Results are returned normalised
 *     (z = 1, or (0, 1, 0) for the identity).
 *   - Every function returns a zkc_status; 0 = OK.  No C++ exception crosses the ABI.  There is
 *     NO CPU fallback: without a CUDA device every compute call returns ZKC_ERR_CUDA.
 *   - `*_dev` variants take device pointers (data already resident in HBM); the plain variants
 *     take host pointers and include the host<->device copies.
 *   - Tier B comes in two forms.  The STEP API (zkc_prove_begin ... zkc_prove_end) keeps randomness and the
 *     Fiat-Shamir transcript on the caller's side of the boundary: every step takes the challenges and the
 *     scalars the caller drew and returns points / evaluations for the caller's own TranscriptWrite.
 *     zkc_prove is a convenience driver over the same steps that owns a transcript and an RNG restated
 *     from upstream (seed and kinds chosen through zkc_prove_opts).
 *   - One zkc_ctx per GPU; calls on one ctx are serialised internally (thread-safe per ctx).
 */
#ifndef ZKCERT_CUDA_H
#define ZKCERT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[4]; } zkc_fr;
typedef struct { uint64_t l[4]; } zkc_fq;
typedef struct { zkc_fq x, y; } zkc_g1_affine;
typedef struct { zkc_fq x, y, z; } zkc_g1;

typedef struct zkc_ctx zkc_ctx;
typedef struct zkc_domain zkc_domain;
typedef struct zkc_srs zkc_srs;
typedef struct zkc_pk zkc_pk;

typedef enum {
  ZKC_OK = 0,
  ZKC_ERR_BAD_ARG = 1,
  ZKC_ERR_CUDA = 2,
  ZKC_ERR_OOM = 3,
  /* mirrors of halo2_proofs::plonk::Error */
  ZKC_ERR_INVALID_INSTANCES = 10,
  ZKC_ERR_CONSTRAINT_SYSTEM_FAILURE = 11, /* e.g. lookup input not in table */
  ZKC_ERR_BOUNDS_FAILURE = 12,
  ZKC_ERR_OPENING = 13,
  ZKC_ERR_SYNTHESIS = 14,
  ZKC_ERR_NOT_ENOUGH_ROWS = 15,
  ZKC_ERR_TRANSCRIPT = 16
} zkc_status;

/* ---- context ------------------------------------------------------------------------------ */
int zkc_ctx_create(int device, zkc_ctx** out);
void zkc_ctx_destroy(zkc_ctx* ctx);
const char* zkc_last_error(const zkc_ctx* ctx);
/* external != 0: adopt an externally owned cudaStream_t (e.g. torch's current stream; the handle 0 is
 * the legacy default stream).  external == 0: go back to the ctx's own non-blocking stream. */
int zkc_ctx_set_stream(zkc_ctx* ctx, void* cuda_stream, int external);
int zkc_ctx_sync(zkc_ctx* ctx);
/* zkc_prove runs work that is off the Fiat-Shamir critical path (coefficient forms, extended cosets) on a second
 * stream underneath the latency-bound MSM phases.  on = 0 serialises everything on one stream (used when timing
 * individual kernels); results are identical either way. */
int zkc_ctx_set_overlap(zkc_ctx* ctx, int on);
/* Debug / sweep overrides on a live ctx: "msm_c", "msm_c_pre", "msm_T", "ntt_two_pass_max", "msm_accum_occ" (0 = library default),
 * "stage_min_bytes" (-1 = default), "team_poison", "team_commit_by_column" (team commitments dealt by column instead of by point
 * range), "no_program_factoring" (keys loaded while set keep their gate programs as parsed).  The ZKC_* environment
 * variables of the same names are read once, when the ctx is created; nothing reads the environment on the proving path. */
int zkc_ctx_set_tunable(zkc_ctx* ctx, const char* name, int64_t value);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
uint64_t zkc_ctx_launch_count(const zkc_ctx* ctx);
const char* zkc_version(void);
/* Per-phase CUDA-event timers on the ctx stream (the NVTX-range equivalent).  While
 * enabled every internal phase (msm.accum, ntt.pass, quotient, ...) is bracketed by events;
 * zkc_profile_report synchronises, then writes a JSON object {"phase": {"ms": total, "n": count}, ...}
 * into buf (truncated to cap) and clears the accumulators. */
int zkc_profile_enable(zkc_ctx* ctx, int on);
int zkc_profile_report(zkc_ctx* ctx, char* buf, size_t cap);
/* same records as a timeline: JSON [["phase", start_ms, duration_ms], ...] relative to the first phase */
int zkc_profile_timeline(zkc_ctx* ctx, char* buf, size_t cap);

/* ---- device memory helpers (so non-CUDA hosts can keep columns resident) ------------------- */
int zkc_dev_alloc(zkc_ctx* ctx, size_t bytes, void** dptr);
int zkc_dev_free(zkc_ctx* ctx, void* dptr);
int zkc_h2d(zkc_ctx* ctx, void* dptr, const void* hptr, size_t bytes);
int zkc_d2h(zkc_ctx* ctx, void* hptr, const void* dptr, size_t bytes);

/* ---- field vectors: halo2curves Fr/Fq arithmetic ---------------------------- */
typedef enum {
  ZKC_OP_ADD = 0, ZKC_OP_SUB = 1, ZKC_OP_MUL = 2, ZKC_OP_INV = 3 /* batch_invert; 0 -> 0 */,
  ZKC_OP_FROM_CANONICAL = 4, ZKC_OP_TO_CANONICAL = 5, ZKC_OP_NEG = 6, ZKC_OP_SQUARE = 7 /* Field::square */
} zkc_vec_op;
/* field: 0 = Fr, 1 = Fq.  out[i] = a[i] (op) b[i]; b may be NULL for unary ops.  Device pointers. */
int zkc_field_vec_op_dev(zkc_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n);

/* poly::batch_invert_assigned: out[i] = num[i] / den[i] with den = 0 -> 0; device pointers. */
int zkc_batch_invert_assigned_dev(zkc_ctx* ctx, const zkc_fr* num, const zkc_fr* den, zkc_fr* out, size_t n);

/* out[i] = first * base^i (arithmetic::powers; used for omega^i / DELTA^c tables at keygen).  Device pointer. */
int zkc_fr_powers_dev(zkc_ctx* ctx, zkc_fr* out_dev, size_t n, const zkc_fr* base, const zkc_fr* first);

/* ---- best_fft (halo2_proofs::arithmetic::best_fft) -------------------------- */
/* In place, natural order in and out: a[j] <- sum_i a[i] * omega^(i*j).  `omega` must have order
 * 2^log_n.  Host buffer variant (drop-in) and device-resident batched variant (ncols contiguous
 * columns of 2^log_n elements). */
int zkc_fft_fr(zkc_ctx* ctx, zkc_fr* a, const zkc_fr* omega, uint32_t log_n);
int zkc_fft_fr_dev(zkc_ctx* ctx, zkc_fr* a_dev, const zkc_fr* omega, uint32_t log_n, uint32_t ncols);

/* ---- EvaluationDomain (halo2_proofs::poly::EvaluationDomain) ----------------- */
typedef struct {
  uint32_t k, extended_k, j;
  zkc_fr omega, omega_inv, extended_omega, extended_omega_inv, g_coset, g_coset_inv;
  zkc_fr ifft_divisor, extended_ifft_divisor;
} zkc_domain_info;
/* EvaluationDomain::new(j, k); zeta_choice selects Fr::ZETA (0 = halo2curves' constant; SURVEY OPEN-4). */
int zkc_domain_create(zkc_ctx* ctx, uint32_t j, uint32_t k, int zeta_choice, zkc_domain** out);
void zkc_domain_free(zkc_domain* dom);
int zkc_domain_get_info(const zkc_domain* dom, zkc_domain_info* out);
/* host-buffer drop-ins */
int zkc_lagrange_to_coeff(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a /* n, in place */);
int zkc_coeff_to_lagrange(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a /* n, in place */);
int zkc_coeff_to_extended(zkc_ctx* ctx, const zkc_domain* dom, const zkc_fr* coeffs /* n */, zkc_fr* out /* 2^extended_k */);
int zkc_extended_to_coeff(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a /* 2^extended_k in place; tail zeroed */);
/* device-resident, batched over ncols contiguous columns */
int zkc_lagrange_to_coeff_dev(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a_dev, uint32_t ncols);
int zkc_coeff_to_lagrange_dev(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a_dev, uint32_t ncols);
int zkc_coeff_to_extended_dev(zkc_ctx* ctx, const zkc_domain* dom, const zkc_fr* coeffs_dev, zkc_fr* out_dev, uint32_t ncols);
int zkc_extended_to_coeff_dev(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a_dev, uint32_t ncols);
int zkc_divide_by_vanishing_dev(zkc_ctx* ctx, const zkc_domain* dom, zkc_fr* a_dev);
/* coeff_to_extended by RESIDUE CLASS (the split team proving uses): the extended coset zeta * <w_ext> is the union of
 * the 2^(extended_k - k) cosets (zeta * w_ext^c) * <omega>; class c of a column is ONE size-n transform and holds the rows
 * c + 2^(extended_k - k) * m of zkc_coeff_to_extended's output at out[c * n + m] (class-major).  Transforms classes [c0, c1) of
 * ncols columns (input stride n, output stride 2^extended_k); the other classes of `out_dev` are left untouched.
 * zkc_extended_classes_to_natural_dev reorders one class-major column into zkc_coeff_to_extended's row order. */
int zkc_coeff_to_extended_classes_dev(zkc_ctx* ctx, const zkc_domain* dom, const zkc_fr* coeffs_dev, zkc_fr* out_dev, uint32_t ncols,
                                      uint32_t c0, uint32_t c1);
int zkc_extended_classes_to_natural_dev(zkc_ctx* ctx, const zkc_domain* dom, const zkc_fr* class_major_dev, zkc_fr* natural_dev);

/* ---- best_multiexp / ParamsKZG ------------------------------------------ */
/* best_multiexp(coeffs, bases) -> G1 (normalised Jacobian).  Host pointers. */
int zkc_msm_g1(zkc_ctx* ctx, const zkc_fr* scalars, const zkc_g1_affine* bases, size_t n, zkc_g1* out);
/* device-resident variant: `ncols` scalar columns of n against the same n bases; out[ncols] on host. */
int zkc_msm_g1_dev(zkc_ctx* ctx, const zkc_fr* scalars_dev, const zkc_g1_affine* bases_dev, size_t n, uint32_t ncols, zkc_g1* out);

/* ParamsKZG: uploads g and g_lagrange once; g2 / s_g2 stay with the host verifier. */
typedef enum { ZKC_BASIS_COEFF = 0, ZKC_BASIS_LAGRANGE = 1 } zkc_basis;
int zkc_srs_load(zkc_ctx* ctx, uint32_t k, const zkc_g1_affine* g, const zkc_g1_affine* g_lagrange, zkc_srs** out);
/* ParamsKZG::setup(k, rng) with the secret handed in (gen_srs draws it from ChaCha20 zero seed,
 * /root/reference/src/helpers.rs:210): g[i] = [s^i]G, g_lagrange[i] = [l_i(s)]G, built on the device. */
int zkc_srs_setup(zkc_ctx* ctx, uint32_t k, const zkc_fr* s, zkc_srs** out);
void zkc_srs_free(zkc_srs* srs);
int zkc_srs_get(zkc_ctx* ctx, const zkc_srs* srs, int basis, zkc_g1_affine* out /* n, host */);
/* ParamsKZG::commit / commit_lagrange: poly has len <= n entries. */
int zkc_commit(zkc_ctx* ctx, const zkc_srs* srs, int basis, const zkc_fr* poly, size_t len, zkc_g1* out);
int zkc_commit_dev(zkc_ctx* ctx, const zkc_srs* srs, int basis, const zkc_fr* polys_dev, size_t len, uint32_t ncols, zkc_g1* out);

uint32_t zkc_srs_k(const zkc_srs* srs);
/* Sum of n normalised Jacobian points on the host: the epilogue of a point-range-sharded MSM across GPUs
 * (each rank runs zkc_msm_g1_dev on its slice of bases/scalars, the 96-byte partials are all-gathered). */
int zkc_g1_sum(const zkc_g1* pts, size_t n, zkc_g1* out);

/* ---- ProvingKey + create_proof (halo2_proofs::plonk::{ProvingKey, create_proof}) ---
 * The reference reaches this through snark-verifier-sdk gen_snark_shplonk
 * (/root/reference/src/helpers.rs:233,299; src/bin/cli.rs:320,343,369,462,519).
 *
 * zkc_pk_load takes what `ProvingKey<G1Affine>` holds: the constraint system (`pk.vk.cs`, serialised as
 * documented in halo2-zkcert_b200/circuit.py), the fixed columns and the permutation sigma columns
 * in the Lagrange basis (`pk.fixed_values`, `pk.permutation.permutations`), and `vk.transcript_repr`.
 * Coefficient forms, extended cosets, l0 / l_last / l_active_row and the vk commitments are rebuilt
 * on the device. */
int zkc_pk_load(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const zkc_fr* fixed /* num_fixed * n */,
                const zkc_fr* sigma /* n_perm_columns * n */, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out);
void zkc_pk_free(zkc_pk* pk);
/* vk.fixed_commitments / vk.permutation.commitments (host buffers sized num_fixed / n_perm_columns) */
int zkc_pk_get_commitments(zkc_ctx* ctx, const zkc_pk* pk, zkc_g1_affine* fixed_out, zkc_g1_affine* sigma_out);
/* out[8] = k, extended_k, cs.degree(), blinding_factors, #permutation sets, #lookups, #fixed, #permutation columns */
int zkc_pk_info(const zkc_pk* pk, uint32_t* out);

/* ---- keygen: what gen_pk does around the commitments (/root/reference/src/helpers.rs:213,265) --------------
 * permutation::keygen::Assembly: replays the copy constraints (4 x u32 each: left column, left row, right column, right row;
 * columns index cs.permutation's column list) in order and returns the cycle structure, mapping[col * n + row] = col' * n + row'
 * (sigma_col[row] = DELTA^col' * omega^row').  Host only.  Out-of-range cells: ZKC_ERR_BOUNDS_FAILURE. */
int zkc_keygen_permutation_mapping(uint32_t k, uint32_t num_columns, const uint32_t* copies, size_t num_copies, uint64_t* mapping_out);
/* ConstraintSystem::compress_selectors: activations = num_selectors columns of n bytes (0 / 1), max_degrees[s] = largest gate
 * degree the selector multiplies (0 = in no gate), max_degree = cs.degree().  Greedy upstream order.  combination_of[s] = fixed
 * column the selector lands in, root_of[s] = the value j >= 1 that column holds on the selector's rows, combination_len[s] = L:
 * the substitution is q * prod_{i = 1..L, i != j} (i - q).  columns_out: up to num_selectors columns of n u32.  Host only. */
int zkc_keygen_compress_selectors(uint32_t k, uint32_t num_selectors, const uint8_t* activations, const uint32_t* max_degrees,
                                  uint32_t max_degree, uint32_t* combination_of, uint32_t* root_of, uint32_t* combination_len,
                                  uint32_t* columns_out, uint32_t* num_combinations);
/* keygen_pk from what synthesis leaves behind: the fixed columns (selectors already substituted) and the copy constraints.
 * The permutation is assembled on the host, the sigma columns are built on the device, the rest is zkc_pk_load. */
int zkc_keygen_pk(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const zkc_fr* fixed /* num_fixed * n */,
                  const uint32_t* copies, size_t num_copies, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out);
/* pk.permutation.permutations (the sigma columns, Lagrange basis) back on the host: n_perm_columns * n */
int zkc_pk_get_sigma(zkc_ctx* ctx, const zkc_pk* pk, zkc_fr* sigma_out);

/* ---- ProvingKey files: `*.pk` as snark-verifier-sdk's gen_pk writes and read_pk reads them with
 * SerdeFormat::RawBytesUnchecked (/root/reference/src/bin/cli.rs:247,312,335,362,455).  Layout as recalled from
 * ProvingKey::write (SURVEY OPEN-8, unpinned): see csrc/host/keygen.cpp.  Offsets of every section: */
typedef struct {
  uint64_t k_off, num_fixed_off, fixed_commitments_off, perm_commitments_off, selectors_off, l0_off, l_last_off, l_active_row_off,
           fixed_values_off, fixed_polys_off, fixed_cosets_off, perm_values_off, perm_polys_off, perm_cosets_off, total;
} zkc_pk_file_layout_t;
int zkc_pk_file_layout(uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm, uint32_t num_selectors, zkc_pk_file_layout_t* out);
int zkc_pk_file_write_headers(uint8_t* file, size_t cap, uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm,
                              uint32_t num_selectors, int be);
int zkc_pk_file_check(const uint8_t* file, size_t len, uint32_t k, uint32_t extended_k, uint32_t num_fixed, uint32_t num_perm,
                      uint32_t num_selectors, int be /* -1 = detect */, int* be_out);
int zkc_fr_column_is_canonical(const zkc_fr* col, size_t n);
/* ProvingKey::write: every section of the resident key into `out` (size from zkc_pk_file_size).  selectors: num_selectors
 * bit-packed activation columns (vk.selectors), may be NULL with num_selectors = 0.  be: 1 = big-endian counts (this revision). */
size_t zkc_pk_file_size(const zkc_pk* pk, uint32_t num_selectors);
int zkc_pk_write(zkc_ctx* ctx, const zkc_pk* pk, const uint8_t* selectors, uint32_t num_selectors, int be, uint8_t* out, size_t cap);
/* ProvingKey::read: fixed_values and permutations stream from the file into the device, everything else is rebuilt there.
 * format: 0 = RawBytes (scalars checked canonical, commitments checked on the curve and against the rebuilt ones),
 * 1 = RawBytesUnchecked (trusted).  The constraint system comes from the circuit, as upstream (`read::<_, ConcreteCircuit>`). */
int zkc_pk_read(zkc_ctx* ctx, const zkc_srs* srs, const uint8_t* cs_blob, size_t cs_len, const uint8_t* file, size_t file_len,
                uint32_t num_selectors, int format, const zkc_fr* transcript_repr, int zeta_choice, zkc_pk** out);

typedef struct {
  int transcript;       /* 0 = Blake2bWrite<_, _, Challenge255>, 1 = Keccak256Write (halo2_proofs::transcript);
                           2 = snark-verifier EvmTranscript (gen_evm_proof_shplonk, cli.rs:519): Keccak-256 sponge over
                               big-endian words, 64-byte uncompressed points in the proof;
                           3 = snark-verifier PoseidonTranscript<_, NativeLoader, _> with POSEIDON_SPEC (T=3, RATE=2, R_F=8,
                               R_P=57) — what gen_snark_shplonk instantiates (helpers.rs:233,299; SURVEY OPEN-7) */
  int multiopen;        /* 0 = ProverSHPLONK, 1 = ProverGWC */
  int advice_blinding;  /* SURVEY OPEN-1: 0 = axiom (last row := 1, no draws), 1 = PSE (unusable rows random) */
  int blind_draws;      /* SURVEY OPEN-2: 1 = one Fr::random per commitment for the (unused) KZG blind */
  int point_format;     /* SURVEY OPEN-5: 0 = y-sign in bit 7; 1 = y-sign in bit 6, identity flag in bit 7 */
  int rng_kind;         /* 0 = rand_chacha::ChaCha20Rng, 1 = rand::rngs::StdRng (ChaCha12; what the SDK's `StdRng` is — OPEN-6) */
  uint8_t rng_seed[32]; /* SeedableRng::from_seed; Fr::random takes 16 keystream words, fill_bytes(32) eight */
  int lookup_fill;      /* SURVEY OPEN-9: 0 = PSE permute_expression_pair (leftover table values go to the repeated rows from the
                           last one backwards), 1 = the axiom fork's rayon variant (ascending rows) */
  int random_poly;      /* SURVEY OPEN-3: 0 = n serial Fr::random draws from the caller's rng; 1 = per-thread ChaCha20Rng instances
                           seeded by rng.fill_bytes, chunks of n / threads coefficients (thread-count dependent upstream) */
  uint32_t random_poly_threads; /* rayon::current_num_threads() of the machine being reproduced (random_poly = 1 only) */
} zkc_prove_opts;

/* create_proof(params, pk, &[circuit], &[instances], rng, &mut transcript) for ONE circuit.
 * `advice`: num_advice columns of n assigned cells (Montgomery; what WitnessCollection holds after
 * batch_invert_assigned), host memory or, with advice_on_device != 0, device memory.  Rows past
 * the usable range are overwritten by the blinding policy.  instances[c] has instance_lens[c] values.
 * The proof bytes (commitments compressed to 32 B, evaluations 32 B) are written to proof_out. */
int zkc_prove(zkc_ctx* ctx, const zkc_pk* pk, const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances,
              const size_t* instance_lens, size_t num_instance_columns /* != cs.num_instance_columns -> InvalidInstances */,
              const zkc_prove_opts* opts, uint8_t* proof_out, size_t proof_cap, size_t* proof_len);

/* Compact witness columns: kind 0 = zkc_fr[n] (Montgomery), 1 = bit-packed (1 bit per cell, LSB first,
 * ceil(n/8) bytes), 2 = uint8[n], 3 = uint16[n], 4 = uint64[n] (canonical small integers).  Host pointers.  The proof is
 * byte-identical to zkc_prove on the expanded columns; the H2D volume of a SHA256-bit witness drops ~250x. */
typedef struct { int kind; const void* data; } zkc_advice_column;
int zkc_prove_compact(zkc_ctx* ctx, const zkc_pk* pk, const zkc_advice_column* cols /* num_advice */, const zkc_fr* const* instances,
                      const size_t* instance_lens, size_t num_instance_columns, const zkc_prove_opts* opts, uint8_t* proof_out,
                      size_t proof_cap, size_t* proof_len);

/* ---- create_proof, STEP API (Tier B) -------------------------------------------------------------------------
 * One call per prover round of plonk::create_proof (halo2_proofs 0.2.0 @4b42325 src/plonk/prover.rs; reached from
 * /root/reference/src/helpers.rs:233,299 through gen_snark_shplonk).  The caller keeps its own `TranscriptWrite` (Blake2b,
 * Keccak, snark-verifier's Poseidon / EVM transcripts - whatever it instantiates) and its own RNG: it writes the returned
 * points / scalars to the transcript, squeezes the challenges and hands them to the next step together with the scalars it
 * drew where upstream draws from `rng`.  Nothing here hashes or draws; given the same inputs the library is deterministic.
 * Order is fixed: begin, lookups, products, vanishing, quotient, evals, then open_shplonk_h + open_shplonk_w or open_gwc, end.
 * A step called out of order returns ZKC_ERR_BAD_ARG; after an error only zkc_prove_end is valid.  One session per ctx at
 * a time.  Points are returned affine, identity = (0, 0) (upstream's write_point then fails on the caller's side).
 * Team proving: every rank runs the same sequence with identical arguments. */
typedef struct zkc_prover zkc_prover;
/* the vanishing argument's random polynomial (n coefficients), described so that it can be produced on the device */
typedef struct {
  int kind;               /* 0 = `scalars`: n host scalars the caller drew (Fr::random, upstream order)
                             1 = the Fr::random draws of ChaCha20Rng / StdRng ::from_seed(seed) that start at keystream word
                                 `first_word` (16 words per draw)
                             2 = chunk j of `chunk_len` coefficients = the Fr::random stream of ChaCha20Rng::from_seed(seeds + 32 j)
                                 (SURVEY OPEN-3: the caller drew the seeds with rng.fill_bytes) */
  const zkc_fr* scalars;
  uint8_t seed[32]; int rng_kind; uint64_t first_word;
  const uint8_t* seeds; uint32_t nseeds; uint64_t chunk_len;
} zkc_random_poly;
/* Round 1.  advice / instances as for zkc_prove.  advice_tails = NULL: axiom policy (last row := 1; SURVEY OPEN-1); else
 * num_advice * (blinding_factors + 1) scalars for the unusable rows of each column (PSE policy, drawn column by column).
 * early_random (optional): the random polynomial, if the caller can already describe it (a seeded RNG whose position at the
 * vanishing argument is known) - it is then generated and committed underneath the first rounds. */
int zkc_prove_begin(zkc_ctx* ctx, const zkc_pk* pk, const zkc_fr* advice, int advice_on_device, const zkc_fr* const* instances,
                    const size_t* instance_lens, size_t num_instance_columns, const zkc_fr* advice_tails,
                    const zkc_random_poly* early_random, zkc_prover** out, zkc_g1_affine* advice_commitments /* num_advice */);
/* Round 2 (after theta).  tails: per lookup, blinding_factors + 1 scalars for A' then as many for S'.  out: A'_0, S'_0, A'_1, ... */
int zkc_prove_lookups(zkc_prover* p, const zkc_fr* theta, const zkc_fr* tails, int lookup_fill, zkc_g1_affine* out /* 2 * #lookups */);
/* Round 3 (after beta, gamma).  tails: blinding_factors scalars per permutation set, then per lookup product (upstream draw
 * order).  out: the permutation product commitments, then the lookup product commitments. */
int zkc_prove_products(zkc_prover* p, const zkc_fr* beta, const zkc_fr* gamma, const zkc_fr* tails, zkc_g1_affine* out /* #sets + #lookups */);
/* Round 4a: commitment to the random polynomial (random = NULL if it was given to zkc_prove_begin). */
int zkc_prove_vanishing(zkc_prover* p, const zkc_random_poly* random, zkc_g1_affine* out /* 1 */);
/* Round 4b (after y): h(X) on the extended coset, divided by X^n - 1, split into degree - 1 pieces.  out: their commitments. */
int zkc_prove_quotient(zkc_prover* p, const zkc_fr* y, zkc_g1_affine* out /* cs.degree() - 1 */);
/* Round 5 (after x): every evaluation, in the order upstream writes them to the transcript (advice, fixed, random, sigma,
 * permutation products, lookups).  *count receives the number (also when cap is too small: ZKC_ERR_BAD_ARG). */
int zkc_prove_evals(zkc_prover* p, const zkc_fr* x, zkc_fr* out, size_t cap, size_t* count);
/* Multiopen, ProverSHPLONK: y and v are squeezed back to back; the caller writes *out, squeezes u, and calls _w. */
int zkc_prove_open_shplonk_h(zkc_prover* p, const zkc_fr* y, const zkc_fr* v, zkc_g1_affine* out /* 1 */);
int zkc_prove_open_shplonk_w(zkc_prover* p, const zkc_fr* u, zkc_g1_affine* out /* 1 */);
/* Multiopen, ProverGWC: one witness commitment per distinct evaluation point, first-appearance order. */
int zkc_prove_open_gwc(zkc_prover* p, const zkc_fr* v, zkc_g1_affine* out, size_t cap, size_t* count);
void zkc_prove_end(zkc_prover* p);

/* host-side helpers mirrored for cross-checking a caller's RNG: Fr::random stream of ChaCha20Rng / StdRng ::from_seed(seed)
 * starting at draw `skip`; rand_core's SeedableRng::seed_from_u64. */
int zkc_rng_fr_random(const uint8_t seed[32], int rng_kind, uint64_t skip, zkc_fr* out, size_t count);
void zkc_seed_from_u64(uint64_t state, uint8_t seed[32]);
/* the transcripts' host hashes: kind 0 = Blake2b-512 with a 16-byte personalisation (NULL = none; halo2 uses
 * "Halo2-Transcript"), 64 bytes out; kind 1 = Keccak-256, 32 bytes out.  No device work. */
int zkc_host_hash(int kind, const uint8_t* personal16, const uint8_t* data, size_t len, uint8_t* out);
/* host field inversion (Montgomery in / out), for cross-checks: field 0 = Fr, 1 = Fq; which 0 = the host driver's fe_inv (binary
 * extended Euclid), 1 = the Fermat ladder.  Returns non-zero on a non-canonical input. */
int zkc_host_fe_inv(int field, int which, const uint64_t* in, uint64_t* out, size_t n);
/* Grain-generated Poseidon parameters of transcript kind 3 (canonical little-endian): 65 x 3 round constants, 3 x 3 MDS. */
int zkc_poseidon_spec(zkc_fr* constants, zkc_fr* mds);

/* ---- verify_proof (host only, no device) ---------------------------------------------------------------------------------
 * plonk::verify_proof::<KZGCommitmentScheme<Bn256>, VerifierSHPLONK / VerifierGWC, Challenge255, TranscriptRead, SingleStrategy>:
 * the self-check snark-verifier-sdk's gen_snark_shplonk runs after create_proof (/root/reference/src/helpers.rs:233,299) and what
 * the reference's tests assert (src/tests/x509_aggregation.rs:64-105).  `cs_blob` as for zkc_pk_load; fixed / sigma commitments
 * as returned by zkc_pk_get_commitments (= the VerifyingKey); g1_gen = params.g[0], g2 / s_g2 = params.g2() / params.s_g2()
 * (halo2curves G2Affine: x.c0, x.c1, y.c0, y.c1, Montgomery limbs).  Only opts->transcript, multiopen and point_format are read.
 * *ok = 1 iff the proof is accepted (a malformed proof is a rejected proof: return value ZKC_OK, *ok = 0). */
typedef struct { zkc_fq x_c0, x_c1, y_c0, y_c1; } zkc_g2_affine;
int zkc_verify(const uint8_t* cs_blob, size_t cs_len, const zkc_g1_affine* fixed_comm, const zkc_g1_affine* sigma_comm,
               const zkc_fr* transcript_repr, const zkc_g1_affine* g1_gen, const zkc_g2_affine* g2, const zkc_g2_affine* s_g2,
               const zkc_fr* const* instances, const size_t* instance_lens, size_t num_instance_columns, const uint8_t* proof,
               size_t proof_len, const zkc_prove_opts* opts, int* ok);
/* G2 pieces of ParamsKZG::setup and the pairing behind the final check: generator of G2, [scalar]P, prod e(g1s[i], g2s[i]) == 1 */
int zkc_g2_generator(zkc_g2_affine* out);
int zkc_g2_mul(const zkc_g2_affine* p, const zkc_fr* scalar, zkc_g2_affine* out);
int zkc_pairing_check(const zkc_g1_affine* g1s, const zkc_g2_affine* g2s, size_t npairs, int* is_one);
/* Test hook for the device program optimiser (csrc/host/cs.h optimize_program: a selector shared by a run of constraints is
 * taken out of the run): the Horner fold acc = acc * mult + e_i of the gate expressions over given query values (one zkc_fr
 * per advice / fixed / instance query, Montgomery), with the stream as parsed (factored = 0) or as the device runs it
 * (factored = 1; *groups = number of factored runs).  Both give the same field element.  Host only. */
int zkc_host_fold_gates(const uint8_t* cs_blob, size_t cs_len, const zkc_fr* advice_q, const zkc_fr* fixed_q, const zkc_fr* instance_q,
                        const zkc_fr* mult, int factored, zkc_fr* out, uint32_t* groups);

/* ---- ParamsKZG files (host only) ------------------------------------------------------------------------------------------
 * `kzg_bn254_{k}.srs` as ParamsKZG::write / read lay it out in SerdeFormat::RawBytes[Unchecked] (SURVEY OPEN-8):
 * u32 k (LE) | g[n] | g_lagrange[n] | g2 | s_g2, points as raw Montgomery limbs (64 B / 128 B).  The arrays read here go to
 * zkc_srs_load (device) and zkc_verify (g2, s_g2).  `checked` != 0 validates every point (RawBytes), 0 trusts the file. */
size_t zkc_params_size(uint32_t k);
int zkc_params_write(uint32_t k, const zkc_g1_affine* g, const zkc_g1_affine* g_lagrange, const zkc_g2_affine* g2, const zkc_g2_affine* s_g2,
                     uint8_t* out, size_t cap);
int zkc_params_read(const uint8_t* in, size_t len, int checked, uint32_t* k, zkc_g1_affine* g, zkc_g1_affine* g_lagrange, zkc_g2_affine* g2,
                    zkc_g2_affine* s_g2);

/* ---- team proving: ONE create_proof over the GPUs of a node ---------------------------------------------
 * One process (and one zkc_ctx) per GPU.  After zkc_team_init every rank calls zkc_srs_*, zkc_pk_load and zkc_prove with
 * IDENTICAL arguments; the library partitions the device work (MSM by point range — the split halo2's best_multiexp makes
 * across rayon threads —, column transforms by column, h(X) by extended-row block) and exchanges results with NCCL on the
 * ctx stream.  Every rank returns the same proof bytes, identical to the single-GPU proof.  The id comes from
 * zkc_team_unique_id on one rank and reaches the others through the host's own channel (torch.distributed, MPI, a file).
 * zkc_team_emulate(world) runs all `world` shards on this one GPU sequentially with the collectives elided (testing). */
#define ZKC_TEAM_ID_BYTES 128
int zkc_team_unique_id(uint8_t id[ZKC_TEAM_ID_BYTES]);
int zkc_team_init(zkc_ctx* ctx, int rank, int world, const uint8_t id[ZKC_TEAM_ID_BYTES]);
int zkc_team_emulate(zkc_ctx* ctx, int world);
int zkc_team_leave(zkc_ctx* ctx);
int zkc_team_info(const zkc_ctx* ctx, int* rank, int* world, int* emulated);
/* the partition arithmetic (host only): contiguous share [lo, hi) of `total` items for `rank`; and the residue classes
 * [c0, c1) of the extended coset (class c = rows c + 2^(extended_k - k) * m, a coset of the size-n subgroup) that `rank`
 * transforms and evaluates: the classes its share of the num_classes * 2^k evaluated rows touches (num_classes = degree - 1:
 * h(X) has fewer than (degree - 1) n coefficients, so that many classes determine it) */
int zkc_team_shard_range(uint64_t total, int world, int rank, uint64_t* lo, uint64_t* hi);
int zkc_team_classes(uint32_t k, uint32_t num_classes, int world, int rank, uint32_t* c0, uint32_t* c1);

#ifdef __cplusplus
}
#endif
#endif /* ZKCERT_CUDA_H */
