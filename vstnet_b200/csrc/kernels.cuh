// kernels.cuh — internal launch interfaces shared by the translation units of libvstb200.
#pragma once
#include <algorithm>
#include "common.cuh"

namespace vst {

enum ConvEpi {
    EPI_RELU = 0,       // out = relu(conv + bias)
    EPI_NONE = 1,       // out = conv + bias
    EPI_ADD = 2,        // out = res + (conv + bias)                     forward coupling  (RevResNet.py:103)
    EPI_SUB = 3,        // out = res - (conv + bias)                     inverse coupling  (RevResNet.py:110-111)
    EPI_ADD_SQZ = 4,    // out = squeeze(res) + (conv + bias)            stride-2 forward  (RevResNet.py:100-103)
    EPI_SUB_UNSQZ = 5,  // unsqueeze(out) = res - (conv + bias)          stride-2 inverse  (RevResNet.py:112-113)
};

struct ConvArgs {
    const float* in;    // [Cin][Hin][Win]
    const float* w;     // packed [Cin][9][CoutPad]
    const float* bias;  // [CoutPad]
    const float* res;   // coupling operand (may alias out for EPI_ADD / EPI_SUB)
    float* out;
    int Cin, Cout, CoutPad, Hin, Win, Hout, Wout;
    int epi;
};

int conv_cout_pad(int Cout);
int launch_pack_conv_weights(const float* w, const float* b, float* wp, float* bp, int Cin, int Cout, int CoutPad,
                             cudaStream_t st);
int launch_conv3x3_ffma(const ConvArgs& a, int stride, cudaStream_t st);

// tensor-core (tcgen05) path, conv_tc.cu
bool tc_eligible(int Cin, int Cout, int stride);
int tc_tile_n(int Cout);
size_t tc_packed_floats(int Cin, int Cout, int N, int terms);
int launch_pack_tc_weights(const float* w, float* wp, int Cin, int Cout, int N, int terms, cudaStream_t st);
int launch_conv3x3_tc(const ConvArgs& a, int terms, cudaStream_t st);

int launch_space_to_depth(const float* in, float* out, int C, int Hin, int Win, cudaStream_t st);
int launch_depth_to_space(const float* in, float* out, int Cout, int Hin, int Win, cudaStream_t st);
int launch_latent_spread(const float* x1, const float* x2, float* z, int Ch, int h, int w, int L, cudaStream_t st);
int launch_latent_gather(const float* z, float* x1, float* x2, int Ch, int h, int w, int L, cudaStream_t st);

}  // namespace vst
