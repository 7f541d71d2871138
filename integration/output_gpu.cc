// Picture output for the GPU decoder: SURVEY.md 8f-2, "device-resident DPB, D2H only in write_out_picture with cropping".
//
// Compiled in place of the reference's framebuf/output.cc (integration/Makefile) against its unchanged headers.  The two
// entry points the DPB calls (framebuf/output.h:5-6; callers framebuf/dpb.cc:383-385, 957) keep their meaning; the samples
// no longer come from storable_picture::imgY/imgUV -- the GPU binding never fills those -- but from the engine frame of
// the picture, copied device->host when, and only when, the DPB releases the picture for output: one cropped 8-bit copy
// (h264r_frame_download_cropped) that waits for the picture's own wave only, so pictures parsed after it keep
// reconstructing underneath.  Frame pictures only (the engine's supported subset); field output stops with an error.
#include "global.h"
#include "input_parameters.h"
#include "dpb.h"
#include "picture.h"
#include "sets.h"
#include "output.h"

#include "h264recon.h"

#include <unistd.h>
#include <vector>

using namespace vio::h264;

// integration/decoder_gpu.cc
namespace vio { namespace h264 {
h264r_ctx*  gpu_engine();
h264r_frame gpu_frame_of_picture(const storable_picture* p);
void        gpu_picture_freed(const storable_picture* p);
} }

namespace {

std::vector<uint8_t> g_out;          // display rectangle of one picture: Y, Cb, Cr back to back

void put(int fd, const uint8_t* p, size_t n)
{
    if ((ssize_t)n != write(fd, p, n)) error(500, "write_out_picture: error writing to YUV file");
}

// write_out_picture (output.cc:109-227) for 8-bit 4:2:0 frames: the display rectangle, planes back to back
void output_picture(VideoParameters* p_Vid, storable_picture* p, int p_out)
{
    const sps_t& sps = *p_Vid->active_sps;
    if (p->non_existing || p_out == -1) return;
    if (sps.chroma_format_idc != 1 || sps.BitDepthY != 8 || sps.BitDepthC != 8 || !sps.frame_mbs_only_flag)
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    int left = 0, right = 0, top = 0, bottom = 0;                 // luma samples (CropUnitX = CropUnitY = 2)
    if (sps.frame_cropping_flag) {
        left = 2 * (int)sps.frame_crop_left_offset; right  = 2 * (int)sps.frame_crop_right_offset;
        top  = 2 * (int)sps.frame_crop_top_offset;  bottom = 2 * (int)sps.frame_crop_bottom_offset;
    }
    const int w = (int)sps.PicWidthInMbs * 16 - left - right, h = (int)sps.FrameHeightInMbs * 16 - top - bottom;
    const size_t ny = (size_t)w * h, nc = ny / 4;
    g_out.resize(ny + 2 * nc);
    const int rc = h264r_frame_download_cropped(gpu_engine(), gpu_frame_of_picture(p), left, right, top, bottom,
                                                g_out.data(), g_out.data() + ny, g_out.data() + ny + nc, w, w / 2);
    if (rc != H264R_OK) error(500, "h264recon: h264r_frame_download_cropped: %s", h264r_strerror(rc));
    put(p_out, g_out.data(), ny);
    put(p_out, g_out.data() + ny, nc);
    put(p_out, g_out.data() + ny + nc, nc);
}

} // namespace

void write_stored_frame(VideoParameters* p_Vid, pic_t* fs, int p_out)
{
    if (p_Vid->out_buffer->is_used || fs->is_used < 3)
        error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));      // unpaired fields: not in the GPU subset
    if (fs->recovery_frame) p_Vid->recovery_flag = 1;
    if (!p_Vid->non_conforming_stream || p_Vid->recovery_flag) output_picture(p_Vid, fs->frame, p_out);
    fs->is_output = 1;
}

void direct_output(VideoParameters* p_Vid, storable_picture* p, int p_out)
{
    if (p->slice.structure != FRAME) error(500, "h264recon: %s", h264r_strerror(H264R_ERR_UNSUPPORTED));
    output_picture(p_Vid, p, p_out);
    p_Vid->calculate_frame_no(p);
    gpu_picture_freed(p);
    delete p;
}
