// include/mfb/particle_generator.hh - particle::container + particle::lattice_generator: Bravais lattices, the masked sc
// lattice and glass loads
// (include/particle_container.hh:30-189, include/particle_generator.hh:129-290, 292-403, 422-515, 520-590).
// The AoS arrays are filled on the device and copied once into the std::vectors the output plugins read
// (get_particles()), instead of a serial host triple loop per component.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "grid_fft.hh"

namespace mfb_detail {
// [setup] keys of the masked lattice, read through config_file::get_value_safe when the config type has it (the reference's
// config_file does; a host that drives the mirror with its own configuration object gets the reference's defaults)
template <class C, class T, class = void> struct cfg_reader {
    static T get(C&, const char*, const char*, T dflt) { return dflt; }
};
template <class C, class T>
struct cfg_reader<C, T, decltype(void(std::declval<C&>().get_value_safe(std::string(), std::string(), std::declval<T>())))> {
    static T get(C& cf, const char* sec, const char* key, T dflt) { return cf.get_value_safe(std::string(sec), std::string(key), dflt); }
};
}  // namespace mfb_detail

#if defined(__has_include)
#if __has_include(<particle_container.hh>)
#define MFB_MIRROR_REFERENCE_CONTAINER 1
#endif
#endif
#ifdef MFB_MIRROR_REFERENCE_CONTAINER
#include <particle_container.hh>   // the reference's own container (include/particle_container.hh:30-189): what its output plugins read
#endif

namespace particle {
enum lattice { lattice_masked = -2, lattice_glass = -1, lattice_sc = 0, lattice_bcc = 1, lattice_fcc = 2, lattice_rsc = 3 };

#ifndef MFB_MIRROR_REFERENCE_CONTAINER
// stand-alone hosts (no reference headers on the include path): the same members
class container {
public:
    std::vector<float> positions32_, velocities32_, mass32_;
    std::vector<double> positions64_, velocities64_, mass64_;
    std::vector<uint32_t> ids32_;
    std::vector<uint64_t> ids64_;
    bool bhas_individual_masses_ = false;
    container() = default;
    container(const container&) = delete;
    void allocate(size_t nump, bool b64reals, bool b64ids, bool bindividualmasses) {
        bhas_individual_masses_ = bindividualmasses;
        if (b64reals) { positions64_.resize(3 * nump); velocities64_.resize(3 * nump); if (bindividualmasses) mass64_.resize(nump); }
        else { positions32_.resize(3 * nump); velocities32_.resize(3 * nump); if (bindividualmasses) mass32_.resize(nump); }
        if (b64ids) ids64_.resize(nump); else ids32_.resize(nump);
    }
    const void* get_pos32_ptr() const { return positions32_.data(); }
    const void* get_pos64_ptr() const { return positions64_.data(); }
    const void* get_vel32_ptr() const { return velocities32_.data(); }
    const void* get_vel64_ptr() const { return velocities64_.data(); }
    const void* get_ids32_ptr() const { return ids32_.data(); }
    const void* get_ids64_ptr() const { return ids64_.data(); }
    const void* get_mass32_ptr() const { return mass32_.data(); }
    const void* get_mass64_ptr() const { return mass64_.data(); }
    size_t get_local_num_particles() const { return ids32_.empty() ? ids64_.size() : ids32_.size(); }
    size_t get_global_num_particles() const { return get_local_num_particles(); }
};
#endif

template <typename field_t>
class lattice_generator {
    mfb_particles* p_ = nullptr;
    container particles_;
    bool b64reals_, b64ids_, bmass_, downloaded_ = false;
    size_t global_num_particles_ = 0;

public:
    // the 2x2x2 species mask of a masked lattice from [setup] DoBaryons / ParticleMaskType (include/particle_generator.hh:184-205)
    static void particle_type_mask(bool do_baryons, unsigned mask_type, int mask[8]) {
        for (int s = 0; s < 8; ++s) mask[s] = do_baryons ? -1 : 0;
        if (!do_baryons) return;
        mask[0] = 0;
        if (mask_type & 1u) mask[1] = mask[2] = mask[4] = 0;   // edge centres
        if (mask_type & 2u) mask[3] = mask[5] = mask[6] = 0;   // face centres
        mask[7] = 1;                                           // baryon at the cell centre
    }
    template <class config_t>
    lattice_generator(lattice lattice_type, int lattice_index, const bool b64reals, const bool b64ids, const bool bwithmasses,
                      size_t IDoffset, const field_t& field, config_t& cf)
        : b64reals_(b64reals), b64ids_(b64ids), bmass_(bwithmasses) {
        int n[3] = {int(field.n_[0]), int(field.n_[1]), int(field.n_[2])};
        if (lattice_type == lattice_glass)
            throw std::runtime_error("glass load: hand the tiled glass positions (glass_loader::glass_posr, grid units) to the "
                                     "constructor overload that takes them - reading the glass file stays host code");
        if (lattice_type == lattice_masked) {
            int mask[8];
            particle_type_mask(mfb_detail::cfg_reader<config_t, bool>::get(cf, "setup", "DoBaryons", false),
                               unsigned(mfb_detail::cfg_reader<config_t, int>::get(cf, "setup", "ParticleMaskType", 3)), mask);
            mfb_detail::check(mfb_particles_create_masked(&p_, n, lattice_index, mask, b64reals, b64ids, bwithmasses));
        } else {
            mfb_detail::check(mfb_particles_create(&p_, n, int(lattice_type), b64reals, b64ids, bwithmasses));
        }
        mfb_detail::check(mfb_particles_set_ids(p_, (unsigned long long)IDoffset));
        global_num_particles_ = mfb_particles_count(p_);
    }
    // glass load (:262-290): `posr` = glass_loader::glass_posr (positions in grid units, [np][3]), `offset` = glass_loader::offset()
    lattice_generator(const std::vector<double>& posr, size_t offset, const bool b64reals, const bool b64ids, size_t IDoffset,
                      const field_t& field)
        : b64reals_(b64reals), b64ids_(b64ids), bmass_(false) {
        int n[3] = {int(field.n_[0]), int(field.n_[1]), int(field.n_[2])};
        mfb_detail::check(mfb_particles_create_glass(&p_, n, posr.data(), posr.size() / 3, (unsigned long long)offset, b64reals, b64ids));
        mfb_detail::check(mfb_particles_set_ids(p_, (unsigned long long)IDoffset));
        global_num_particles_ = mfb_particles_count(p_);
    }
    ~lattice_generator() { if (p_) mfb_particles_destroy(p_); }
    template <class config_t>
    void set_masses(const lattice, int lattice_index, const real_t munit, const bool, field_t& field, config_t&) {
        mfb_detail::check(mfb_particles_set_masses(p_, lattice_index, munit, field.device_handle()));
        downloaded_ = false;
    }
    template <class config_t>
    void set_positions(const lattice, int lattice_index, int idim, real_t lunit, const bool, field_t& field, config_t&) {
        mfb_detail::check(mfb_particles_set_positions(p_, lattice_index, idim, lunit, field.device_handle()));
        downloaded_ = false;
    }
    template <class config_t>
    void set_velocities(lattice, int lattice_index, int idim, const bool, field_t& field, config_t&) {
        mfb_detail::check(mfb_particles_set_velocities(p_, lattice_index, idim, field.device_handle()));
        downloaded_ = false;
    }
    const container& get_particles() {
        if (!downloaded_) {
            particles_.allocate(global_num_particles_, b64reals_, b64ids_, bmass_);
            void* pos = b64reals_ ? (void*)particles_.positions64_.data() : (void*)particles_.positions32_.data();
            void* vel = b64reals_ ? (void*)particles_.velocities64_.data() : (void*)particles_.velocities32_.data();
            void* ids = b64ids_ ? (void*)particles_.ids64_.data() : (void*)particles_.ids32_.data();
            void* mas = bmass_ ? (b64reals_ ? (void*)particles_.mass64_.data() : (void*)particles_.mass32_.data()) : nullptr;
            mfb_detail::check(mfb_particles_download(p_, pos, vel, ids, mas));
            downloaded_ = true;
        }
        return particles_;
    }
    // The container chunk by chunk, for particle loads that do not fit the host (3LPT 2048^3: 275 GB): `consume(chunk, first)` is
    // called with a container holding particles [first, first + chunk.get_local_num_particles()) - the same members an output
    // plugin reads - while the device -> host copy of the next chunk is in flight; two chunk containers exist at any time.
    // Extension beyond the reference's interface (its write_particle_data takes the whole container).
    template <class consumer_t> void for_each_particle_chunk(size_t chunk_particles, consumer_t consume) {
        container buf[2];
        const size_t np = global_num_particles_;
        auto issue = [&](int slot, size_t first) {
            const size_t cnt = std::min(chunk_particles, np - first);
            container& c = buf[slot];
            c.allocate(cnt, b64reals_, b64ids_, bmass_);
            void* pos = b64reals_ ? (void*)c.positions64_.data() : (void*)c.positions32_.data();
            void* vel = b64reals_ ? (void*)c.velocities64_.data() : (void*)c.velocities32_.data();
            void* ids = b64ids_ ? (void*)c.ids64_.data() : (void*)c.ids32_.data();
            mfb_detail::check(mfb_particles_stream_out_range(p_, MFB_ATTR_POS, first, cnt, pos));
            mfb_detail::check(mfb_particles_stream_out_range(p_, MFB_ATTR_VEL, first, cnt, vel));
            mfb_detail::check(mfb_particles_stream_out_range(p_, MFB_ATTR_IDS, first, cnt, ids));
            if (bmass_) mfb_detail::check(mfb_particles_stream_out_range(p_, MFB_ATTR_MASS, first, cnt, b64reals_ ? (void*)c.mass64_.data() : (void*)c.mass32_.data()));
        };
        if (np == 0 || chunk_particles == 0) return;
        issue(0, 0);
        int slot = 0;
        for (size_t first = 0; first < np; first += chunk_particles, slot ^= 1) {
            mfb_detail::check(mfb_particles_stream_wait(p_));                  // chunk `first` has landed
            if (first + chunk_particles < np) issue(slot ^ 1, first + chunk_particles);
            consume(static_cast<const container&>(buf[slot]), first);
        }
    }
};
}  // namespace particle
