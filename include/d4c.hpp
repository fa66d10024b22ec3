// Source-compatible replacement for /root/reference/include/d4c.hpp (D4COption :16-20, class D4C
// :23-36) over libworldb200.so.
#ifndef WORLD_CLASS_D4C_HPP
#define WORLD_CLASS_D4C_HPP

#include "harvest.hpp"

namespace world_class
{

typedef struct D4COption{
	double threshold;

	D4COption() : threshold(world::kThreshold) {}
} D4COption;


class D4C
{

public:

	D4C(int fs) : handle_(nullptr) { wb_throw_if(wb_d4c_create(fs, nullptr, &handle_), "wb_d4c_create"); }
	D4C(int fs, const D4COption &option) : handle_(nullptr)
	{
		WbD4COption o;
		o.threshold = option.threshold;
		wb_throw_if(wb_d4c_create(fs, &o, &handle_), "wb_d4c_create");
	}
	~D4C() { wb_d4c_destroy(handle_); }
	D4C(D4C &&other) noexcept : handle_(other.handle_) { other.handle_ = nullptr; }
	D4C(const D4C &) = delete;
	D4C &operator=(const D4C &) = delete;

	void compute(
		const double *x, int x_length,
		const double *temporal_positions, const double *f0, int f0_length,
		int fft_size, double **aperiodicity
	)
	{ wb_throw_if(wb_d4c_compute(handle_, x, x_length, temporal_positions, f0, f0_length, fft_size, aperiodicity), "wb_d4c_compute"); }

private:

	wb_d4c_t *handle_;
};

} // end namespace world_class

#endif
