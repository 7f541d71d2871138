// Picture output for the GPU decoder: SURVEY.md 8f-2, "device-resident DPB, D2H only in write_out_picture with cropping".
//
// Compiled in place of the reference's framebuf/output.cc (integration/Makefile) against its unchanged headers.  The two
// entry points the DPB calls (framebuf/output.h:5-6; callers framebuf/dpb.cc:383-385, 957) keep their meaning; the samples
// no longer come from storable_picture::imgY/imgUV -- the GPU binding never fills those -- but from the engine frame of
// the picture, copied device->host when, and only when, the DPB releases the picture for output: one cropped 8-bit copy
// (h264r_frame_download_cropped) that waits for the picture's own wave only, so pictures parsed after it keep
// reconstructing underneath.  A frame coded as two field pictures lives in the engine as two pictures of half the height:
// both come down and their lines are interleaved here; an unpaired field goes out with an empty (mid-grey) other field,
// as write_unpaired_field does.
#include "global.h"
#include "input_parameters.h"
#include "dpb.h"
#include "picture.h"
#include "sets.h"
#include "output.h"

#include "h264recon.h"

#include <string.h>
#include <unistd.h>
#include <vector>

using namespace vio::h264;

// integration/decoder_gpu.cc
namespace vio { namespace h264 {
h264r_ctx*  gpu_engine_of(const storable_picture* p);
h264r_frame gpu_frame_of_picture(const storable_picture* p);
void        gpu_picture_freed(const storable_picture* p);
bool        gpu_has_picture(const storable_picture* p);
} }

namespace {

std::vector<uint8_t> g_out;          // display rectangle of one picture: Y, Cb, Cr back to back
std::vector<uint8_t> g_field[2];     // the two fields of a frame that was coded as field pictures

void put(int fd, const uint8_t* p, size_t n)
{
    if ((ssize_t)n != write(fd, p, n)) error(500, "write_out_picture: error writing to YUV file");
}

void crop_of(const sps_t& sps, int& left, int& right, int& top, int& bottom)
{
    left = right = top = bottom = 0;                              // luma samples: CropUnitX = 2, CropUnitY = 2 (2 - frame_mbs_only_flag)
    if (sps.frame_cropping_flag) {
        const int uy = 2 * (2 - (int)sps.frame_mbs_only_flag);
        left = 2 * (int)sps.frame_crop_left_offset; right  = 2 * (int)sps.frame_crop_right_offset;
        top  = uy * (int)sps.frame_crop_top_offset; bottom = uy * (int)sps.frame_crop_bottom_offset;
    }
}

// write_out_picture (output.cc:109-227) for 8-bit 4:2:0 frames: the display rectangle, planes back to back
void output_picture(VideoParameters* p_Vid, storable_picture* p, int p_out)
{
    const sps_t& sps = *p_Vid->active_sps;
    if (p->non_existing || p_out == -1) return;
    if (sps.chroma_format_idc != 1 || sps.BitDepthY != 8 || sps.BitDepthC != 8)
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    int left, right, top, bottom;
    crop_of(sps, left, right, top, bottom);
    const int w = (int)sps.PicWidthInMbs * 16 - left - right, h = (int)sps.FrameHeightInMbs * 16 - top - bottom;
    const size_t ny = (size_t)w * h, nc = ny / 4;
    g_out.resize(ny + 2 * nc);
    const int rc = h264r_frame_download_cropped(gpu_engine_of(p), gpu_frame_of_picture(p), left, right, top, bottom,
                                                g_out.data(), g_out.data() + ny, g_out.data() + ny + nc, w, w / 2);
    if (rc != H264R_OK) error(500, "h264recon: h264r_frame_download_cropped: %s", h264r_strerror(rc));
    put(p_out, g_out.data(), ny);
    put(p_out, g_out.data() + ny, nc);
    put(p_out, g_out.data() + ny + nc, nc);
}

// A frame coded as two field pictures: each field is an engine picture of half the height.  Both come down whole and their
// lines are interleaved into the display rectangle -- dpb_combine_field_yuv (framebuf/dpb.cc) on the way out.
void output_field_pair(VideoParameters* p_Vid, storable_picture* top_field, storable_picture* bottom_field, int p_out)
{
    const sps_t& sps = *p_Vid->active_sps;
    if (p_out == -1) return;
    if (sps.chroma_format_idc != 1 || sps.BitDepthY != 8 || sps.BitDepthC != 8)
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    const int W = (int)sps.PicWidthInMbs * 16, H = (int)sps.FrameHeightInMbs * 16, Hf = H / 2;
    const size_t fy = (size_t)W * Hf, fc = fy / 4;
    storable_picture* fields[2] = { top_field, bottom_field };
    for (int k = 0; k < 2; ++k) {
        g_field[k].resize(fy + 2 * fc);
        if (!fields[k]) {                                          // unpaired field: the other one is "empty" (storable_picture::clear, framebuf/picture.cc:129-143)
            memset(g_field[k].data(), 128, fy + 2 * fc);
            continue;
        }
        const int rc = h264r_frame_download(gpu_engine_of(fields[k]), gpu_frame_of_picture(fields[k]), g_field[k].data(), g_field[k].data() + fy,
                                            g_field[k].data() + fy + fc, W, W / 2);
        if (rc != H264R_OK) error(500, "h264recon: h264r_frame_download: %s", h264r_strerror(rc));
    }
    int left, right, top, bottom;
    crop_of(sps, left, right, top, bottom);
    const int w = W - left - right, h = H - top - bottom;
    const size_t ny = (size_t)w * h, nc = ny / 4;
    g_out.resize(ny + 2 * nc);
    for (int y = 0; y < h; ++y) {                                  // frame line top + y = line (top + y) / 2 of field (top + y) & 1
        const int fl = top + y;
        memcpy(g_out.data() + (size_t)y * w, g_field[fl & 1].data() + (size_t)(fl >> 1) * W + left, (size_t)w);
    }
    for (int pl = 0; pl < 2; ++pl)
        for (int y = 0; y < h / 2; ++y) {
            const int fl = top / 2 + y;
            memcpy(g_out.data() + ny + pl * nc + (size_t)y * (w / 2),
                   g_field[fl & 1].data() + fy + pl * fc + (size_t)(fl >> 1) * (W / 2) + left / 2, (size_t)(w / 2));
        }
    put(p_out, g_out.data(), ny + 2 * nc);
}

void release_and_delete(storable_picture*& p)
{
    if (!p) return;
    gpu_picture_freed(p);
    delete p;
    p = nullptr;
}

} // namespace

// write_unpaired_field (output.cc:228-267): a frame store that holds one field only is written with an empty other field.
// The DPB's book-keeping is the reference's (the empty field object, the host-side combination, is_used = 3); the samples
// come from the engine.
static void write_unpaired(VideoParameters* p_Vid, pic_t* fs, int p_out)
{
    const bool have_top = (fs->is_used & 1) != 0;
    storable_picture* p = have_top ? fs->top_field : fs->bottom_field;
    storable_picture*& other = have_top ? fs->bottom_field : fs->top_field;
    other = new storable_picture(p_Vid, have_top ? BOTTOM_FIELD : TOP_FIELD, p->size_x, p->size_y * 2, p->size_x_cr, p->size_y_cr * 2, 1);
    other->clear();
    fs->dpb_combine_field_yuv(p_Vid);
    output_field_pair(p_Vid, have_top ? p : nullptr, have_top ? nullptr : p, p_out);
    fs->is_used = 3;
}

// flush_direct_output (output.cc:269-287): a directly output field that is still waiting for its pair goes out unpaired
static void flush_direct_output(VideoParameters* p_Vid, int p_out)
{
    pic_t* ob = p_Vid->out_buffer;
    if (!ob->is_used) return;
    write_unpaired(p_Vid, ob, p_out);
    if (ob->frame) { delete ob->frame; ob->frame = nullptr; }
    release_and_delete(ob->top_field);
    release_and_delete(ob->bottom_field);
    ob->is_used = 0;
}

void write_stored_frame(VideoParameters* p_Vid, pic_t* fs, int p_out)
{
    flush_direct_output(p_Vid, p_out);
    if (fs->is_used < 3) { write_unpaired(p_Vid, fs, p_out); fs->is_output = 1; return; }
    if (fs->recovery_frame) p_Vid->recovery_flag = 1;
    if (!p_Vid->non_conforming_stream || p_Vid->recovery_flag) {
        // a frame store filled by two field pictures holds them as top_field / bottom_field; its `frame` is the host-side
        // combination the DPB made (dpb_combine_field), which no engine picture stands for
        // (a PAFF stream: the frame form exists on the device if the frame was decoded as one, or was needed as a reference of one)
        if (fs->frame && gpu_has_picture(fs->frame)) output_picture(p_Vid, fs->frame, p_out);
        else output_field_pair(p_Vid, fs->top_field, fs->bottom_field, p_out);
    }
    fs->is_output = 1;
}

// direct_output (output.cc:287-352): pictures that never enter the DPB.  Fields wait in p_Vid->out_buffer for their other half.
void direct_output(VideoParameters* p_Vid, storable_picture* p, int p_out)
{
    pic_t* ob = p_Vid->out_buffer;
    if (p->slice.structure == FRAME) {
        flush_direct_output(p_Vid, p_out);
        output_picture(p_Vid, p, p_out);
        p_Vid->calculate_frame_no(p);
        gpu_picture_freed(p);
        delete p;
        return;
    }
    if (p->slice.structure == TOP_FIELD) {
        if (ob->is_used & 1) flush_direct_output(p_Vid, p_out);
        ob->top_field = p; ob->is_used |= 1;
    } else {
        if (ob->is_used & 2) flush_direct_output(p_Vid, p_out);
        ob->bottom_field = p; ob->is_used |= 2;
    }
    if (ob->is_used == 3) {
        output_field_pair(p_Vid, ob->top_field, ob->bottom_field, p_out);
        p_Vid->calculate_frame_no(p);
        release_and_delete(ob->top_field);
        release_and_delete(ob->bottom_field);
        ob->is_used = 0;
    }
}
