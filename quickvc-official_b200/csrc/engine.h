// engine.h -- internal helpers shared by the translation units of libqvc_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qvc {

int from_series_major(const float* src, int ld, int64_t src_bs, float* dst, int batch, int channels,
                      int frames, bool reverse, cudaStream_t stream);
// (B, C, T) fp32 -> [B][T][C] operand format; frames at or past live[b] (when given) become zero rows
int to_series_major(const float* src, void* dst, int batch, int channels, int frames, int opformat,
                    const int32_t* live, cudaStream_t stream);
int cond_vectors(const float* w, const float* bias, const float* g, int n_embed, int rows, float* out,
                 cudaStream_t stream);
int reflect_row(void* base, int64_t bstride_bytes, int row_bytes, int batch, cudaStream_t stream);

int spk_sequences_per_cluster(int nseq);   // lstm.cu: windows one 8-CTA LSTM cluster carries
void tc_reserve_sms(int n);         // conv_tc.cu: SMs the persistent convolution grids leave free (per host thread)

}  // namespace qvc
