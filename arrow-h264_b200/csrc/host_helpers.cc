// Host-side helpers of the reconstruction boundary that restate what the reference computes per slice header
// (not per sample): implicit bi-prediction weights and the InvLevelScale tables.
#include "h264recon.h"
#include "h264_tables.h"

#include <stdlib.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

extern "C" {

void h264r_implicit_weights(int cur_poc, int poc0, int poc1, int long_term0, int long_term1, int* w0, int* w1)
{
    // restates inter_prediction.cc:112-139 for frame pictures
    int td = clip3(-128, 127, poc1 - poc0);
    if (td == 0 || long_term0 || long_term1) { *w0 = 32; *w1 = 32; return; }
    int tb = clip3(-128, 127, cur_poc - poc0);
    int tx = (16384 + abs(td / 2)) / td;
    int dsf = clip3(-1024, 1023, (tx * tb + 32) >> 6);
    *w1 = dsf >> 2;
    *w0 = 64 - *w1;
    if (*w1 < -64 || *w1 > 128) { *w0 = 32; *w1 = 32; }
}

void h264r_build_level_scale(h264r_slice* s, const int* const q4[6], const int* const q8[2])
{
    // restates Transform::set_quant (transform.cc:265-302): LevelScale = normAdjust * weightScale
    for (int inter = 0; inter < 2; ++inter)
        for (int pl = 0; pl < 3; ++pl)
            for (int k = 0; k < 6; ++k)
                for (int j = 0; j < 4; ++j)
                    for (int i = 0; i < 4; ++i)
                        s->level_scale_4x4[inter][pl][k][j * 4 + i] =
                            (uint16_t)(h264r::norm_adjust_4x4(k, i, j) * q4[inter * 3 + pl][j * 4 + i]);
    for (int inter = 0; inter < 2; ++inter)
        for (int k = 0; k < 6; ++k)
            for (int j = 0; j < 8; ++j)
                for (int i = 0; i < 8; ++i)
                    s->level_scale_8x8[inter][k][j * 8 + i] =
                        (uint16_t)(h264r::norm_adjust_8x8(k, i, j) * q8[inter][j * 8 + i]);
}

} // extern "C"
