// Source-compatible replacement for /root/reference/include/harvest.hpp (public section
// identical: HarvestOption :16-28, class Harvest :31-44); the private section is an opaque
// handle into libworldb200.so.  All work runs on the GPU; bad arguments or a missing GPU throw
// std::runtime_error (the reference has undefined behaviour / void returns).
#ifndef WORLD_CLASS_HARVEST_HPP
#define WORLD_CLASS_HARVEST_HPP

#include <stdexcept>
#include <string>

#include "world_common.hpp"
#include "world_fft.hpp"
#include "worldb200.h"

namespace world_class
{

inline void wb_throw_if(int rc, const char *what) {
	if (rc != WB_OK) throw std::runtime_error(std::string("worldb200: ") + what + " failed (status " + std::to_string(rc) + ")");
}

typedef struct HarvestOption{
	double f0_floor;
	double f0_ceil;
	double frame_period;

	double target_fs;
	double channels_in_octave;

	bool use_cos_table;

	HarvestOption()
		: f0_floor(world::kFloorF0), f0_ceil(world::kCeilF0), frame_period(5)
		, target_fs(8000.), channels_in_octave(40.), use_cos_table(false) {}
	void copy(const HarvestOption& option) { *this = option; }
} HarvestOption;


class Harvest
{

public:

	Harvest(const int fs, const HarvestOption &option) : option_(option), handle_(nullptr)
	{
		WbHarvestOption o;
		o.f0_floor = option.f0_floor; o.f0_ceil = option.f0_ceil; o.frame_period = option.frame_period;
		o.target_fs = option.target_fs; o.channels_in_octave = option.channels_in_octave;
		o.use_cos_table = option.use_cos_table ? 1 : 0;
		wb_throw_if(wb_harvest_create(fs, &o, &handle_), "wb_harvest_create");
	}
	~Harvest() { wb_harvest_destroy(handle_); }
	// the reference is copied by value in test/test.cpp:97 (elided); moving transfers the handle
	Harvest(Harvest &&other) noexcept : option_(other.option_), handle_(other.handle_) { other.handle_ = nullptr; }
	Harvest(const Harvest &) = delete;
	Harvest &operator=(const Harvest &) = delete;

	void compute(
		const double* x, int x_length,double *temporal_positions, double *f0
	)
	{ wb_throw_if(wb_harvest_compute(handle_, x, x_length, temporal_positions, f0), "wb_harvest_compute"); }

	int getSamples(int fs, int x_length, double frame_period) { return wb_harvest_get_samples(fs, x_length, frame_period); }
	int getSamples(int fs, int x_length) { return wb_harvest_get_samples(fs, x_length, option_.frame_period); }

private:

	HarvestOption option_;
	wb_harvest_t *handle_;
};

} // end namespace world_class

#endif
