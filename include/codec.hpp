// Source-compatible replacement for /root/reference/include/codec.hpp:23-88 over libworldb200.so.
// The reference's functions have C linkage and return void; failures here print to stderr
// through the library and leave the outputs untouched.
#ifndef WORLD_CODEC_HPP
#define WORLD_CODEC_HPP

#include "macrodefinitions.hpp"
#include "worldb200.h"

static inline int GetNumberOfAperiodicities(int fs) { return wb_get_number_of_aperiodicities(fs); }

static inline void CodeAperiodicity(const double * const *aperiodicity, int f0_length,
					  int fs, int fft_size, double **coded_aperiodicity)
{ (void)wb_code_aperiodicity(aperiodicity, f0_length, fs, fft_size, coded_aperiodicity); }

static inline void DecodeAperiodicity(const double * const *coded_aperiodicity,
						int f0_length, int fs, int fft_size, double **aperiodicity)
{ (void)wb_decode_aperiodicity(coded_aperiodicity, f0_length, fs, fft_size, aperiodicity); }

static inline void CodeSpectralEnvelope(const double * const *spectrogram, int f0_length,
						  int fs, int fft_size, int number_of_dimensions,
						  double **coded_spectral_envelope)
{ (void)wb_code_spectral_envelope(spectrogram, f0_length, fs, fft_size, number_of_dimensions, coded_spectral_envelope); }

static inline void DecodeSpectralEnvelope(const double * const *coded_spectral_envelope,
							int f0_length, int fs, int fft_size, int number_of_dimensions,
							double **spectrogram)
{ (void)wb_decode_spectral_envelope(coded_spectral_envelope, f0_length, fs, fft_size, number_of_dimensions, spectrogram); }

#endif
