// ops.h -- internal C++ entry points of the kernels (one launch each), used by the handle-level
// orchestration in encoder.cu / llama.cu and re-exported 1:1 through the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/seedb200.h"

namespace sb {

int gemm(const seedb200_gemm_desc& d, cudaStream_t stream);
int attention(const seedb200_attn_desc& d, cudaStream_t stream);
int layernorm(const void* x, int64_t ldx, const void* w, const void* b, void* y, int64_t ldy, int rows, int cols,
              float eps, cudaStream_t stream);
int row_stats(const void* x, int64_t ldx, int rows, int cols, float eps, void* stats, cudaStream_t stream);
int row_stats_from_moments(const void* moments, int rows, int cols, float eps, void* stats, cudaStream_t stream);
int ln_fold_weights(const void* W, int64_t ldw, const void* gamma, const void* beta, const void* bias, int N, int K,
                    void* W_out, void* c_out, void* b_out, cudaStream_t stream);
int rmsnorm(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy, int rows, int cols, float eps,
            cudaStream_t stream);
int patchify(const void* images, int B, void* cols, int kpad, cudaStream_t stream);
int broadcast_rows(const void* src, int src_rows, int cols, void* dst, int64_t ldd, int64_t group_stride_rows,
                   int groups, cudaStream_t stream);
int embedding(const void* table, int64_t ld, const int64_t* ids, int n, int cols, void* out, int64_t ldo,
              int64_t n_rows, cudaStream_t stream);
int vq_argmin(const void* z, const void* codebook, int n, int n_codes, int dim, int mode, int64_t* ids,
              cudaStream_t stream);
// dyn (optional, device): dyn[0] overrides past_len at run time (graph-replayed decode step)
int rope_kv_append_tables(const void* qkv, const int64_t* positions, int B, int S, int H, int D, int past_len,
                          int max_seq, int max_pos, const void* cos_t, const void* sin_t, void* q_out,
                          void* k_cache, void* v_cache, cudaStream_t stream, const int* dyn = nullptr);
int build_rope_tables(void* cos_t, void* sin_t, int max_pos, int D, float base, cudaStream_t stream);
int get_rope_tables(int D, float base, int min_pos, const void** cos_t, const void** sin_t, int* max_pos,
                    cudaStream_t stream);
// elementwise helpers (misc.cu)
int add_rows(const void* a, const void* b, void* out, int rows, int cols, int b_rows, cudaStream_t stream);
int gemv(const void* x, const void* W, int64_t ldw, void* out, const void* residual, const void* norm_w, float eps,
         int M, int N, int K, int mode, cudaStream_t stream, int64_t ldo = 0);   // ldo 0 = dense output rows
// dyn (optional, device): kv_len = dyn[0] + 1 at run time; the grid is then sized for max_seq keys
int decode_attention(const void* q, const void* k_cache, const void* v_cache, void* out, int B, int H, int D,
                     int kv_len, int max_seq, float scale, void* workspace, cudaStream_t stream,
                     const int* dyn = nullptr, int* tickets = nullptr);
int decode_attention_max_splits(int max_seq);
// RoPE of the new token's q / k + KV append + attention in one launch (max_seq <= 2048); dyn as above (past_len = dyn[0])
bool decode_attention_rope_supported(int D, int max_seq);
int decode_attention_rope(const void* qkv, const int64_t* positions, int B, int H, int D, int past_len, int max_seq,
                          int max_pos, const void* cos_t, const void* sin_t, void* k_cache, void* v_cache, void* out,
                          float scale, cudaStream_t stream, const int* dyn = nullptr);
// sampler.cu
struct GenParams { seedb200_sample_params sp; long long eos, pad; };
// gp (host, by value) or gp_dev (device, read at run time); state (device, optional) = {cache length, step, arrive,
// valid steps, flag}: the step index comes from state[1] and the last CTA advances the counters.
int sample(const void* logits, int64_t ld, int B, int V, const GenParams* gp, const GenParams* gp_dev, uint64_t step,
           int* state, int advance_cache, int64_t* tokens, int64_t* out, int64_t out_ld, int* finished,
           cudaStream_t stream);
int image_ids_to_tokens(const int64_t* ids, int n, int64_t shift, int64_t boi, int64_t eoi, int64_t* out,
                        int64_t out_stride, cudaStream_t stream);
int cur_device();

}  // namespace sb
