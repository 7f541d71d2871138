/* TEST INFRASTRUCTURE ONLY.  C API of the CPU restatement (oracle/port_recon.c) -- same shape as the
 * reference harness (oracle/ref_harness.cc: ref_*), so tests drive both through one code path. */
#ifndef PORT_RECON_H_
#define PORT_RECON_H_
#include "h264recon.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct port_dec port_dec;
port_dec* port_open(const h264r_seq_params* sp);
void      port_close(port_dec* d);
int       port_frame_alloc(port_dec* d);
void      port_frame_release(port_dec* d, int id);
void      port_frame_get(port_dec* d, int id, uint8_t* y, uint8_t* cb, uint8_t* cr);
void      port_frame_set(port_dec* d, int id, const uint8_t* y, const uint8_t* cb, const uint8_t* cr);
int       port_reconstruct(port_dec* d, int dst, const h264r_pic_params* pp, int used_for_reference,
                           const h264r_slice* slices, const h264r_mb* mbs, const h264r_mb_motion* motion,
                           const h264r_level* levels, double* sec_decode, double* sec_deblock);
#ifdef __cplusplus
}
#endif
#endif
