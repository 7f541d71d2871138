// T0 of SURVEY.md 8a: the selection of the scaling lists with fall-back rules A and B.
//
// Links Decoder::assign_quant_params of the GPU binding (integration/decoder_gpu.cc) and the REFERENCE's own Transform::init
// (decoder/transform.cc:173-262, compiled unmodified from /root/reference) into one host program and compares, for every
// combination of seq/pic_scaling_matrix_present_flag and of the eight per-list present flags of the SPS and of the PPS
// (2 x 2 x 256 x 256 patterns, UseDefaultScalingMatrix flags and list contents drawn at random), the eight lists both
// select and -- on a sample of the patterns -- every InvLevelScale entry h264r_build_level_scale derives from them against
// the reference's set_quant.  No GPU is touched (libh264recon.so is only linked).  Built and run by tests/test_quant_select.py.
#define private public             // Transform::qmatrix / InvLevelScale* are private members of the reference's class
#include "global.h"
#include "slice.h"
#include "sets.h"
#include "decoder.h"
#undef private

#include "h264recon.h"

#include <stdio.h>
#include <string.h>

namespace vio { namespace h264 { const int* gpu_quant_list(const Decoder* d, int i); } }
using namespace vio::h264;

static uint64_t rng_state = 0x51A1E5ull;
static uint32_t rnd()
{
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
}

int main()
{
    sps_t* sps = new sps_t();
    pps_t* pps = new pps_t();
    sps->chroma_format_idc = 1;
    pps->transform_8x8_mode_flag = 1;
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 16; ++k) { sps->ScalingList4x4[i][k] = 1 + rnd() % 255; pps->ScalingList4x4[i][k] = 1 + rnd() % 255; }
    for (int i = 0; i < 6; ++i) for (int k = 0; k < 64; ++k) { sps->ScalingList8x8[i][k] = 1 + rnd() % 255; pps->ScalingList8x8[i][k] = 1 + rnd() % 255; }
    slice_t* slice = new slice_t;
    slice->active_sps = sps; slice->active_pps = pps;
    Transform* ref = new Transform;
    long patterns = 0, tables = 0;
    for (int sp = 0; sp < 2; ++sp)
        for (int pp = 0; pp < 2; ++pp)
            for (int sf = 0; sf < (sp ? 256 : 1); ++sf)
                for (int pf = 0; pf < (pp ? 256 : 1); ++pf) {
                    sps->seq_scaling_matrix_present_flag = sp; pps->pic_scaling_matrix_present_flag = pp;
                    const uint32_t r = rnd();
                    for (int i = 0; i < 8; ++i) {
                        sps->seq_scaling_list_present_flag[i] = (sf >> i) & 1; pps->pic_scaling_list_present_flag[i] = (pf >> i) & 1;
                        if (i < 6) { sps->UseDefaultScalingMatrix4x4Flag[i] = (r >> i) & 1; pps->UseDefaultScalingMatrix4x4Flag[i] = (r >> (8 + i)) & 1; }
                        else { sps->UseDefaultScalingMatrix8x8Flag[i - 6] = (r >> i) & 1; pps->UseDefaultScalingMatrix8x8Flag[i - 6] = (r >> (8 + i)) & 1; }
                    }
                    ref->init(*slice);                                   // the reference
                    slice->decoder.assign_quant_params(*slice);          // the GPU binding
                    for (int i = 0; i < 8; ++i) {
                        const int* ours = gpu_quant_list(&slice->decoder, i);
                        if (!ours || memcmp(ours, ref->qmatrix[i], sizeof(int) * (i < 6 ? 16 : 64))) {
                            printf("list %d differs: sps %d (flags %02x) pps %d (flags %02x) use-default %04x\n", i, sp, sf, pp, pf, r & 0xFFFF);
                            return 1;
                        }
                    }
                    ++patterns;
                    if (patterns % 97 == 0) {                            // the derived tables (set_quant, transform.cc:265-302)
                        static h264r_slice hs;
                        const int* q4[6]; const int* q8[2];
                        for (int i = 0; i < 6; ++i) q4[i] = gpu_quant_list(&slice->decoder, i);
                        for (int i = 0; i < 2; ++i) q8[i] = gpu_quant_list(&slice->decoder, 6 + i);
                        h264r_build_level_scale(&hs, q4, q8);
                        for (int pl = 0; pl < 3; ++pl) for (int k = 0; k < 6; ++k) for (int j = 0; j < 4; ++j) for (int i = 0; i < 4; ++i)
                            if (hs.level_scale_4x4[0][pl][k][j * 4 + i] != ref->InvLevelScale4x4_Intra[pl][k][j][i] ||
                                hs.level_scale_4x4[1][pl][k][j * 4 + i] != ref->InvLevelScale4x4_Inter[pl][k][j][i]) { printf("4x4 table differs\n"); return 1; }
                        for (int k = 0; k < 6; ++k) for (int j = 0; j < 8; ++j) for (int i = 0; i < 8; ++i)
                            if (hs.level_scale_8x8[0][k][j * 8 + i] != ref->InvLevelScale8x8_Intra[0][k][j][i] ||
                                hs.level_scale_8x8[1][k][j * 8 + i] != ref->InvLevelScale8x8_Inter[0][k][j][i]) { printf("8x8 table differs\n"); return 1; }
                        ++tables;
                    }
                }
    printf("scaling-list selection ok: %ld flag patterns equal to Transform::init, %ld derived table sets equal to set_quant\n", patterns, tables);
    return 0;
}
