// tests/cpp/output_raw_plugin.cc - TEST INFRASTRUCTURE: an output plugin written against the reference's UNCHANGED plugin API
// (include/output_plugin.hh:33-123: virtual interface + self-registering output_plugin_creator_concrete), selected by
// "[output] format = mfb_raw".  It writes what the driver hands it as raw little-endian files next to [output] filename:
//   <filename>.<SPECIES>_<component>.f64   fields [n0][n1][n2]     (raw_particles = no: species as Lagrangian fields, like "generic")
//   <filename>.<SPECIES>.{pos,vel}.f64 / .ids.u64 / .mass.f64       (raw_particles = yes: the particle container, like "gadget")
// With more than one MPI rank every rank writes its slab / its particles to <filename>.r<rank>.*
// so that tests/test_reference_driver.py can read the output of the reference's own executable without HDF5.
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include <output_plugin.hh>

class raw_output_plugin : public output_plugin {
    bool particles_, eulerian_, real64_, ids64_;
    double lunit_, vunit_, munit_ = 1.0;
    std::unique_ptr<output_plugin> units_of_;   // raw_units_of = <plugin>: units, widths and species mode are asked of that plugin
    static std::string species(const cosmo_species& s) { return s == cosmo_species::dm ? "DM" : (s == cosmo_species::baryon ? "BA" : "NU"); }
    static void dump(const std::string& name, const void* p, size_t bytes) {
        FILE* f = std::fopen(name.c_str(), "wb");
        if (!f) throw std::runtime_error("mfb_raw: cannot open " + name);
        if (bytes && std::fwrite(p, 1, bytes, f) != bytes) { std::fclose(f); throw std::runtime_error("mfb_raw: short write to " + name); }
        std::fclose(f);
    }

public:
    explicit raw_output_plugin(config_file& cf, std::unique_ptr<cosmology::calculator>& pcc) : output_plugin(cf, pcc, "raw dump") {
        particles_ = cf_.get_value_safe<bool>("output", "raw_particles", false);
        eulerian_ = cf_.get_value_safe<bool>("output", "raw_eulerian", false);
        // units and widths of the particle data: what the driver asks every output plugin for (defaults: generic's)
        real64_ = cf_.get_value_safe<bool>("output", "raw_real64", true);
        ids64_ = cf_.get_value_safe<bool>("output", "raw_ids64", true);
        lunit_ = cf_.get_value_safe<double>("output", "raw_position_unit", 1.0);
        vunit_ = cf_.get_value_safe<double>("output", "raw_velocity_unit", 1.0);
        // stand in for another output plugin of the reference: everything the driver asks an output plugin for (units, 32 / 64
        // bit, particles or fields) is answered by an instance of THAT plugin, made by the reference's own factory; only the
        // bytes go to raw files instead of HDF5
        const std::string other = cf_.get_value_safe<std::string>("output", "raw_units_of", "");
        if (!other.empty()) {
            auto it = get_output_plugin_map().find(other);
            if (it == get_output_plugin_map().end()) throw std::runtime_error("mfb_raw: unknown output plugin '" + other + "'");
            units_of_ = it->second->create(cf, pcc);
            real64_ = units_of_->has_64bit_reals(); ids64_ = units_of_->has_64bit_ids();
            lunit_ = units_of_->position_unit(); vunit_ = units_of_->velocity_unit(); munit_ = units_of_->mass_unit();
            particles_ = units_of_->write_species_as(cosmo_species::dm) == output_type::particles;
            eulerian_ = units_of_->write_species_as(cosmo_species::dm) == output_type::field_eulerian;
        }   // Eulerian fields: the semiclassical (PPT) branch, like generic_out_eulerian
        if (CONFIG::MPI_task_size > 1) fname_ += ".r" + std::to_string(CONFIG::MPI_task_rank);   // one set of files per rank: its slab
    }
    output_type write_species_as(const cosmo_species&) const {
        return eulerian_ ? output_type::field_eulerian : (particles_ ? output_type::particles : output_type::field_lagrangian);
    }
    bool has_64bit_reals() const { return real64_; }
    bool has_64bit_ids() const { return ids64_; }
    real_t position_unit() const { return lunit_; }
    real_t velocity_unit() const { return vunit_; }
    real_t mass_unit() const { return munit_; }

    void write_grid_data(const Grid_FFT<real_t>& g, const cosmo_species& s, const fluid_component& c) {
        static const char* comp[] = {"delta", "vx", "vy", "vz", "dx", "dy", "dz", "mass"};   // fluid_component (include/general.hh:60-69)
        if (int(c) < 0 || int(c) > 7) throw std::runtime_error("mfb_raw: unknown fluid component");
        std::vector<double> v(g.size(0) * g.size(1) * g.size(2));   // the local planes (the whole grid on one rank)
        for (size_t i = 0; i < g.size(0); ++i)
            for (size_t j = 0; j < g.size(1); ++j)
                for (size_t k = 0; k < g.size(2); ++k) v[(i * g.size(1) + j) * g.size(2) + k] = g.relem(i, j, k);
        dump(fname_ + "." + species(s) + "_" + comp[int(c)] + ".f64", v.data(), v.size() * sizeof(double));
    }
    void write_particle_data(const particle::container& pc, const cosmo_species& s, double) {
        const std::string base = fname_ + "." + species(s);
        const size_t np = pc.get_local_num_particles();
        if (real64_) {
            dump(base + ".pos.f64", pc.get_pos64_ptr(), 3 * np * sizeof(double));
            dump(base + ".vel.f64", pc.get_vel64_ptr(), 3 * np * sizeof(double));
            if (pc.bhas_individual_masses_) dump(base + ".mass.f64", pc.get_mass64_ptr(), np * sizeof(double));
        } else {
            dump(base + ".pos.f32", pc.get_pos32_ptr(), 3 * np * sizeof(float));
            dump(base + ".vel.f32", pc.get_vel32_ptr(), 3 * np * sizeof(float));
            if (pc.bhas_individual_masses_) dump(base + ".mass.f32", pc.get_mass32_ptr(), np * sizeof(float));
        }
        if (ids64_) dump(base + ".ids.u64", pc.get_ids64_ptr(), np * sizeof(uint64_t));
        else dump(base + ".ids.u32", pc.get_ids32_ptr(), np * sizeof(uint32_t));
    }
};

namespace {
output_plugin_creator_concrete<raw_output_plugin> creator_raw("mfb_raw");
}
