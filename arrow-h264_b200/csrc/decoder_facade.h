#ifndef H264R_DECODER_FACADE_H_
#define H264R_DECODER_FACADE_H_
#include "h264recon.h"
#endif
