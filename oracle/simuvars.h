/* TEST INFRASTRUCTURE — CPU oracle for `scssim simuvars`. Not shipped, never
 * linked into or called from the product (scssim_b200/); only tests/ and bench legs that time the CPU
 * side use it.
 *
 * Restates, with the reference's own string surgery and its own integer types, what
 *   Genome::loadAbers        /root/reference/lib/genome/Genome.cpp:35-165
 *   SNPOnChr::readSNPs, SNP  /root/reference/lib/snp/snp.cpp:13-36,147-203
 *   Genome::saveSequence     /root/reference/lib/genome/Genome.cpp:329-386
 *   Genome::generateSegment  /root/reference/lib/genome/Genome.cpp:388-691
 * compute. The only random draws are `randomInteger(0, ploidy)` = libc rand() scaled
 * (lib/mydefine/MyDefine.cpp:285-292); `main` never calls srand() on the simuvars branch
 * (src/scssim.cpp:33-38), so the reference's output is a deterministic function of its inputs and
 * glibc's default seed (1). The oracle therefore calls libc srand(seed)/rand() itself, and parity
 * is pinned by byte-comparing its FASTA with that of the compiled reference (oracle/_ref/bin/scssim
 * simuvars) in tests/test_simuvars_oracle.py, plus the committed fixture tests/golden/simuvars_small.
 */
#pragma once
#include <map>
#include <string>
#include <vector>

namespace orc {

struct SvCnv { long spos, epos; float cn, mcn; };
struct SvSnp { long long pos; char nuc; };
struct SvSnv { long pos; char alt; bool het; };
struct SvIns { long pos; std::string seq; bool het; };
struct SvDel { long pos; int len; bool het; };

struct SimuVars {
    int ploidy = 2;                     /* Config.cpp:35 */
    unsigned seed = 1;                  /* glibc's state when srand() was never called */
    std::string err;
    std::vector<std::string> chroms;    /* FASTA index order, "chr"/"chrom" stripped (Fasta.cpp:57-68) */
    std::vector<std::string> seqs;      /* as in the file (case kept) */
    std::map<std::string, std::vector<SvCnv>> cnvs;
    std::map<std::string, std::vector<SvSnp>> snps;
    std::map<std::string, std::vector<SvSnv>> snvs;
    std::map<std::string, std::vector<SvIns>> inss;
    std::map<std::string, std::vector<SvDel>> dels;
    long nCnv = 0, nSnv = 0, nIns = 0, nDel = 0, nSnp = 0;

    bool load_fasta(const std::string& path);
    bool load_vars(const std::string& path);    /* empty path: nothing to load */
    bool load_snps(const std::string& path);
    /* whole output FASTA (names `<chr>_<k>_<chromLen>`, 100 bases per line) */
    bool run(std::string& out);

  private:
    void segment(std::vector<std::string>& hap, const std::string& chr, const std::string& chrSeq, long s, long e, int CN, int mCN);
};

}  // namespace orc
