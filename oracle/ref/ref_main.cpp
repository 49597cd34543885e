/*
 * ref_main.cpp — a two-subcommand front door onto the reference's UNMODIFIED sources.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/): compiles /root/reference/src/phylocsf++build_tracks.hpp and
 * phylocsf++score_msa.hpp where they lie, with oracle/ref/gsl/ standing in for the absent GSL. It is the reference's
 * own main() (src/phylocsf++.cpp:19-61) minus the three annotation tools, which need libBigWig (fetched over the network
 * by the reference's CMake, CMakeLists.txt:122-131) and are outside the hot path.
 * Output: oracle/_ref/phylocsf_ref — `phylocsf_ref build-tracks|score-msa <the reference's own options>`.
 */
#define __STDC_FORMAT_MACROS
#include <inttypes.h>

#include "arg_parse.hpp"

#ifdef ENABLE_OPENMP
    #include <omp.h>
#else
    int omp_get_max_threads() { return 1; }
    int omp_get_thread_num() { return 0; }
#endif

#include "phylocsf++build_tracks.hpp"
#include "phylocsf++score_msa.hpp"

int main(int argc, char **argv)
{
    if (argc >= 2 && strcmp(argv[1], "build-tracks") == 0)
        return main_build_tracks(argc - 1, argv + 1);
    if (argc >= 2 && strcmp(argv[1], "score-msa") == 0)
        return main_score_msa(argc - 1, argv + 1);
    fprintf(stderr, "usage: %s build-tracks|score-msa [the reference's options]\n", argv[0]);
    return 2;
}
