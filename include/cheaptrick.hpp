// Source-compatible replacement for /root/reference/include/cheaptrick.hpp (CheapTrickOption
// :14-20, class CheapTrick :23-38) over libworldb200.so.
#ifndef WORLD_CLASS_CHEAPTRICK_HPP
#define WORLD_CLASS_CHEAPTRICK_HPP

#include "harvest.hpp"

namespace world_class
{

typedef struct CheapTrickOption{
	double q1;
	double f0_floor;
	int fft_size;

	CheapTrickOption() : q1(-0.15), f0_floor(world::kFloorF0), fft_size(0) {}
} CheapTrickOption;


class CheapTrick
{

public:

	CheapTrick(int fs) : handle_(nullptr) { wb_throw_if(wb_cheaptrick_create(fs, nullptr, &handle_), "wb_cheaptrick_create"); }
	CheapTrick(int fs, const CheapTrickOption &option) : handle_(nullptr)
	{
		WbCheapTrickOption o;
		o.q1 = option.q1; o.f0_floor = option.f0_floor; o.fft_size = option.fft_size;
		wb_throw_if(wb_cheaptrick_create(fs, &o, &handle_), "wb_cheaptrick_create");
	}
	~CheapTrick() { wb_cheaptrick_destroy(handle_); }
	CheapTrick(CheapTrick &&other) noexcept : handle_(other.handle_) { other.handle_ = nullptr; }
	CheapTrick(const CheapTrick &) = delete;
	CheapTrick &operator=(const CheapTrick &) = delete;

	void compute(
		const double *x, int x_length, const double *temporal_positions,
		const double *f0, int f0_length, double **spectrogram
	)
	{ wb_throw_if(wb_cheaptrick_compute(handle_, x, x_length, temporal_positions, f0, f0_length, spectrogram), "wb_cheaptrick_compute"); }

	int getFFTSizeForCheapTrick(int fs, double f0_floor) { return wb_cheaptrick_get_fft_size(fs, f0_floor); }

	double getF0FloorForCheapTrick(int fs, int fft_size) { return wb_cheaptrick_get_f0_floor(fs, fft_size); }

private:

	wb_cheaptrick_t *handle_;
};

} // end namespace world_class

#endif
