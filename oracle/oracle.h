/* oracle.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the reference ggml CPU backend's arithmetic for the hot path
 * (quantised matmul, activation quantisers, flash attention, glue ops).  Every function
 * cites the reference file:line it follows (paths relative to /root/reference/llama.cpp).
 * The restatement is pinned against the reference itself (oracle/_ref/libggml-*.so, built
 * from the reference sources by oracle/Makefile) in tests/test_oracle_pin.py, and against
 * the committed golden vectors in tests/golden/ that were generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may
 * use this library.
 */
#ifndef B200_ORACLE_H
#define B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml type ids we care about (ggml/include/ggml.h:350-390) */
enum {
    ORC_TYPE_F32 = 0, ORC_TYPE_F16 = 1, ORC_TYPE_Q4_0 = 2, ORC_TYPE_Q8_0 = 8,
    ORC_TYPE_Q4_K = 12, ORC_TYPE_Q5_K = 13, ORC_TYPE_Q6_K = 14, ORC_TYPE_Q8_K = 15,
};

/* fp16 <-> fp32, IEEE round-to-nearest-even (== F16C _cvtss_sh / _cvtsh_ss) */
uint16_t orc_f32_to_f16(float f);
float    orc_f16_to_f32(uint16_t h);

size_t orc_row_size(int type, int64_t k);      /* bytes of one row of k elements */
int    orc_block_elems(int type);               /* 32 or 256 */

/* activation quantisers exactly as the reference CPU backend (x86 AVX2 build) computes them */
void orc_quantize_row_q8_0(const float *x, void *y, int64_t k);   /* ggml-cpu-quants.c:808-860 (RNE, id=127/amax) */
void orc_quantize_row_q8_K(const float *x, void *y, int64_t k);   /* ggml-quants.c:2479-2513 */
void orc_quantize_row_q4_0(const float *x, void *y, int64_t k);   /* ggml-quants.c:35-70 (KV store f32->q4_0) */

/* dequantisers: ggml-quants.c:255 (q4_0), :349 (q8_0), :1280 (q4_K), :1482 (q5_K), :1690 (q6_K) */
void orc_dequantize_row(int type, const void *x, float *y, int64_t k);

/* per-block exact integer sums of the CPU vec_dot.
 * For each weight block b writes P[b] and M[b] (M is 0 for types without mins):
 *   q4_0/q8_0 vs q8_0 : P = sum w*a                       (ggml-cpu-quants.c:1912, :3663)
 *   q4_K/q5_K vs q8_K : P = sum_j sc_j sum q*a, M = sum bsums_j * m_{j/2}   (:7267-7322, :7326)
 *   q6_K vs q8_K      : P = sum_j scales_j sum (q-32)*a   (:8148) */
void orc_block_sums(int type, const void *w, const void *act, int64_t k, int32_t *P, int32_t *M);

/* float result of one row dot (scalar-order combination of the exact integer sums) */
float orc_vec_dot(int type, const void *w, const void *act, int64_t k);

/* dst[m*N + n] = W[n,:] . x[m,:]   W: N rows quantised `type` with K elems, x: f32 [M,K]
 * follows ggml_compute_forward_mul_mat (ggml-cpu.c:8708-8900): quantise src1 rows to the
 * weight's vec_dot_type, then one vec_dot per (row, col). */
void orc_mul_mat(int type, const void *W, const float *x, float *dst, int64_t N, int64_t K, int64_t M);

/* mul_mat_id (ggml-cpu.c:8902-...): as [K,N,n_expert]; b f32 [K, nb1(n_used or 1), n_tok];
 * ids i32 [n_used, n_tok]; dst f32 [N, n_used, n_tok] */
void orc_mul_mat_id(int type, const void *as, const float *b, const int32_t *ids, float *dst,
                    int64_t N, int64_t K, int64_t n_expert, int64_t n_used, int64_t n_tok, int64_t b_ne1);

/* glue (ggml-cpu.c): rms_norm :6920-, rope :10573-10800, soft_max :9930-, silu, get_rows... */
void orc_rms_norm(const float *x, float *y, int64_t ncols, int64_t nrows, float eps);
/* rope mode 0 (norm) or 2 (neox); x/y [ne0, n_head, n_tok] contiguous */
void orc_rope(const float *x, float *y, const int32_t *pos, const float *freq_factors,
              int64_t ne0, int64_t n_head, int64_t n_tok, int n_dims, int mode, int n_ctx_orig,
              float freq_base, float freq_scale, float ext_factor, float attn_factor,
              float beta_fast, float beta_slow);
void orc_soft_max(const float *x, const uint16_t *mask_f16, float *y, int64_t ncols, int64_t nrows, float scale);
void orc_silu_mul(const float *gate, const float *up, float *y, int64_t n);

/* flash_attn_ext (ggml-cpu.c:12221-12434): q f32 [D, n_q, H] contiguous, k/v [D, n_kv, Hkv]
 * rows of `type_k`/`type_v` (f16, q8_0, q4_0) with byte strides; mask f16 [n_kv, n_q_pad];
 * dst f32 [D, H, n_q]. */
void orc_flash_attn_ext(const float *q, const void *k, const void *v, const uint16_t *mask, float *dst,
                        int64_t D, int64_t n_q, int64_t H, int64_t n_kv, int64_t Hkv,
                        int type_k, int type_v,
                        size_t k_nb1, size_t k_nb2, size_t v_nb1, size_t v_nb2, size_t mask_nb1,
                        float scale, float logit_softcap);

#ifdef __cplusplus
}
#endif
#endif
