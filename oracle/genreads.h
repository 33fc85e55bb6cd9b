/* TEST INFRASTRUCTURE — CPU oracle (see draws.h header). Array-based restatement of the
 * reference's `scssim genreads` stages, in the reference's own sequential (-t 1) order:
 *   splitToFrags     /root/reference/lib/genome/Genome.cpp:753-782, lib/fragment/Fragment.cpp:40-50
 *   createPrimers    lib/malbac/Malbac.cpp:36-103
 *   setPrimers       Malbac.cpp:236-283, poissRand lib/mydefine/MyDefine.cpp:69-80
 *   amplify          Malbac.cpp:173-201, Fragment.cpp:52-137, lib/amplicon/Amplicon.cpp:156-240
 *   list order       Amplicon.cpp:574-585 (prepend in a task), Malbac.cpp:105-141 (append task lists)
 *   getSequence      Amplicon.cpp:255-382
 *   setReadCounts    Malbac.cpp:370-408, randIndx_hp MyDefine.cpp:203-272, getGCFactor Profile.cpp:1503
 *   yieldReads       Malbac.cpp:410-460, Amplicon.cpp:402-565
 * Sequences are never stored per fragment: a template base is read from the genome text through
 * the same reverse/complement maps the reference applies to its copies.
 */
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "draws.h"
#include "profile.h"

namespace orc {

struct Params {
    long primers = 100000; double gamma = 1e-9; double coverage = 5; int isize = 260; bool paired = true;
    double ber = 3.4e-4; int ampMin = 1000, ampMax = 2000, fragSize = 1000;
    int fragMin = 10000, fragMax = 100000;
};

struct Frag { int seq; int64_t start0; int len; int strand; int primers; };
struct Err { uint32_t pos; uint8_t alt; };
struct Amp {
    uint32_t tmpl;   /* semi: fragment index; full: index into the semi list */
    uint32_t spos, len, gc; uint16_t primers;
    uint32_t err_off, err_n;   /* into Sim::errs */
};

struct Sim {
    Params P; Profile prof; Draws* D = nullptr;
    std::vector<std::string> names; std::vector<std::string> seqs;   /* upper-cased */
    std::vector<Frag> frags; std::vector<Amp> semis, fulls; std::vector<Err> errs;
    std::vector<long> primerCount; uint64_t totalPrimers = 0;
    std::vector<double> gcFactors, weights; std::vector<uint32_t> readNumbers; uint64_t reads = 0;
    std::vector<uint64_t> semiBatchEnd, fullBatchEnd;   /* list sizes after each batch */
    std::string fq1, fq2;   /* PE: _1/_2 ; SE: fq1 */
    uint64_t nRecords = 0;
    std::string err;

    bool load_fasta(const std::string& path);
    void set_genome(const std::vector<std::string>& n, const std::vector<std::string>& s) { names = n; seqs = s; }
    void split_to_frags();
    void amplify();
    void set_read_counts();
    void yield_reads();
    void run() { split_to_frags(); amplify(); yield_reads(); }

    /* template base accessors */
    char frag_tmpl(const Frag& f, uint32_t i) const;
    void semi_window(const Amp& s, std::string& out) const;   /* errored window, un-reversed */
    void semi_tmpl(const Amp& s, std::string& out) const;     /* complement(reverse(window)) */
    void full_sequence(const Amp& a, std::string& out) const;

  private:
    void create_primers();
    void set_primers(bool onlyFrags, int round);
    long poiss_rand(double lambda);
    int take_primer(const char* s8);
    template <class GetBase> void amplify_template(GetBase tb, uint32_t length, int primerNum, bool fromFragment,
                                                   uint32_t tmplIdx, std::vector<Amp>& out);
    void amplify_frags(int pass);
    void amplify_semis(int cycle);
    static void append_record(std::string& out, long ampIdx, int fragCount, const char* suffix,
                              const std::string& s, const std::string& q);
};

}  // namespace orc
