// Source-compatible replacement for /root/reference/tools/parameterio.hpp:24-117 over libworldb200.so
// (csrc/wb_io.cu).  Same names, argument meaning and return values (the Read* functions return 1 on
// success and 0 on failure, like the reference); the files are byte-identical to the reference's.
#ifndef WORLD_PARAMETERIO_HPP
#define WORLD_PARAMETERIO_HPP

#include <stdio.h>

#include "worldb200.h"

static inline void WriteF0(const char *filename, int f0_length, double frame_period,
				 const double *temporal_positions, const double *f0, int text_flag) {
	if (wb_write_f0(filename, f0_length, frame_period, temporal_positions, f0, text_flag) != WB_OK)
		printf("File cannot be opened.\n");
}

static inline int ReadF0(const char *filename, double *temporal_positions, double *f0) {
	return wb_read_f0(filename, temporal_positions, f0) == WB_OK ? 1 : 0;
}

static inline double GetHeaderInformation(const char *filename, const char *parameter) {
	return wb_get_header_information(filename, parameter);
}

static inline void WriteSpectralEnvelope(const char *filename, int fs, int f0_length,
							   double frame_period, int fft_size, int number_of_dimensions,
							   const double * const *spectrogram) {
	if (wb_write_spectral_envelope(filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, spectrogram) != WB_OK)
		printf("File cannot be opened.\n");
}

static inline int ReadSpectralEnvelope(const char *filename, double **spectrogram) {
	return wb_read_spectral_envelope(filename, spectrogram) == WB_OK ? 1 : 0;
}

static inline void WriteAperiodicity(const char *filename, int fs, int f0_length,
						   double frame_period, int fft_size, int number_of_dimensions,
						   const double * const *aperiodicity) {
	if (wb_write_aperiodicity(filename, fs, f0_length, frame_period, fft_size, number_of_dimensions, aperiodicity) != WB_OK)
		printf("File cannot be opened.\n");
}

static inline int ReadAperiodicity(const char *filename, double **aperiodicity) {
	return wb_read_aperiodicity(filename, aperiodicity) == WB_OK ? 1 : 0;
}

#endif
