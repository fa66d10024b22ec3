// Source-compatible replacement for /root/reference/include/synthesis.hpp (class Synthesis
// :29-51) over libworldb200.so.
#ifndef WORLD_CLASS_SYNTHESIS_HPP
#define WORLD_CLASS_SYNTHESIS_HPP

#include "harvest.hpp"

namespace world_class
{

class Synthesis
{

public:

	// fs: sampling frequency, fft_size: FFT size, frame_period: analysis frame period [ms]
	Synthesis(int fs, int fft_size, double frame_period) : handle_(nullptr)
	{ wb_throw_if(wb_synthesis_create(fs, fft_size, frame_period, &handle_), "wb_synthesis_create"); }
	~Synthesis() { wb_synthesis_destroy(handle_); }
	Synthesis(Synthesis &&other) noexcept : handle_(other.handle_) { other.handle_ = nullptr; }
	Synthesis(const Synthesis &) = delete;
	Synthesis &operator=(const Synthesis &) = delete;

	// out (out_length samples, allocated by the caller) is fully overwritten
	void compute(
		const double *f0, int f0_length,
		const double * const *spectrogram, const double * const *aperiodicity,
		int out_length, double *out
	)
	{ wb_throw_if(wb_synthesis_compute(handle_, f0, f0_length, spectrogram, aperiodicity, out_length, out), "wb_synthesis_compute"); }

private:

	wb_synthesis_t *handle_;
};

} // end namespace world_class

#endif
