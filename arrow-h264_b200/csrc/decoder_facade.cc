// placeholder translation unit: the reference-interface mirror (class Decoder) lives in decoder_facade.h
#include "decoder_facade.h"
