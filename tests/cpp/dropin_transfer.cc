// tests/cpp/dropin_transfer.cc - links the reference's UNMODIFIED transfer-function plugin (src/plugins/transfer_eisenstein.cc,
// src/transfer_function_plugin.cc, src/cosmology_parameters.cc) compiled against the drop-in headers, selects it through the
// reference's own plugin factory and tabulates T(k) the way monofonic_b200.pipeline.radial_table does: one compute() call per
// distinct |k|.  Prints "k T_total" pairs.
#include <cstdio>
#include <general.hh>
#include <config_file.hh>
#include <cosmology_parameters.hh>
#include <transfer_function_plugin.hh>

std::map<cosmo_species, std::string> cosmo_species_name = {{cosmo_species::dm, "dm"}, {cosmo_species::baryon, "baryon"},
                                                           {cosmo_species::neutrino, "neutrino"}};
namespace CONFIG {
int MPI_thread_support = -1, MPI_task_rank = 0, MPI_task_size = 1;
bool MPI_ok = false, MPI_threads_ok = false, FFTW_threads_ok = false;
int num_threads = 1;
}  // namespace CONFIG
size_t global_mem_high_mark, local_mem_high_mark;
extern "C" { const char* GIT_TAG = "dropin"; const char* GIT_REV = "dropin"; const char* GIT_BRANCH = "dropin"; }

int main(int argc, char** argv) {
    if (argc < 2) return 1;
    config_file cf(argv[1]);
    cosmology::parameters cosmo(cf);
    auto tf = select_TransferFunction_plugin(cf, cosmo);
    tf->intialise();
    for (double k = 1e-3; k < 20.0; k *= 1.7) std::printf("%.10e %.10e\n", k, tf->compute(k, delta_matter));
    return 0;
}
