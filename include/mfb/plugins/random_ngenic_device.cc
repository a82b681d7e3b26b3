// include/mfb/plugins/random_ngenic_device.cc - a noise plugin for the reference's UNCHANGED RNG plugin API
// (include/random_plugin.hh:26-65: virtual Fill_Grid + self-registering RNG_plugin_creator_concrete), to be dropped next to
// src/plugins/random_ngenic.cc:   [random] generator = NGENIC_GPU   seed = <as for NGENIC>
// Same white noise as the NGENIC plugin for the same seed and grid - RANLUX (gsl_rng_ranlxd1) column streams seeded from the
// same seed table, src/plugins/random_ngenic.cc:42-154 - but generated on the device (mfb_noise_fill_ngenic_ranlxd): no serial
// host fill of N^3/2 modes, no H2D copy of the grid, no GSL.  Works on slabs (every rank fills its own k-slab).
#include <general.hh>
#include <random_plugin.hh>
#include <config_file.hh>

class RNG_ngenic_device : public RNG_plugin {
    unsigned long long seed_;

public:
    explicit RNG_ngenic_device(config_file& cf) : RNG_plugin(cf) { seed_ = (unsigned long long)cf.get_value<long>("random", "seed"); }
    virtual ~RNG_ngenic_device() {}
    bool isMultiscale() const { return false; }
    void Fill_Grid(Grid_FFT<real_t>& g) {
        g.FourierTransformForward(false);   // the fill is a k-space fill (random_ngenic.cc:86-87)
        mfb_detail::check(mfb_noise_fill_ngenic_ranlxd(g.device_handle(), seed_));
    }
};

namespace {
RNG_plugin_creator_concrete<RNG_ngenic_device> creator_ngenic_device("NGENIC_GPU");
}
