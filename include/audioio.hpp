// Source-compatible replacement for /root/reference/tools/audioio.hpp:16-58 over libworldb200.so
// (wb_wavwrite / wb_get_audio_length / wb_wavread, csrc/wb_io.cu).  Same names, argument meaning and
// return values; where the reference prints a message and returns, the library returns a status that
// these shims report on stdout like the reference does.
#ifndef WORLD_AUDIOIO_HPP
#define WORLD_AUDIOIO_HPP

#include <stdio.h>

#include "worldb200.h"

static inline void wavwrite(const double *x, int x_length, int fs, int nbit, const char *filename) {
	if (wb_wavwrite(x, x_length, fs, nbit, filename) != WB_OK) printf("File cannot be opened.\n");
}

static inline int GetAudioLength(const char *filename) { return wb_get_audio_length(filename); }

static inline void wavread(const char *filename, int *fs, int *nbit, double *x) {
	if (wb_wavread(filename, fs, nbit, x) != WB_OK) printf("File not found.\n");
}

#endif
