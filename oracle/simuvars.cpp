/* TEST INFRASTRUCTURE — see simuvars.h. Plain std::string surgery, the reference's integer types. */
#include "simuvars.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace orc {

static std::string strip_chr(std::string s) {   /* abbrOfChr MyDefine.cpp:310-323 == aberOfChr snp.cpp:131-145 */
    size_t i = s.find("chrom");
    if (i == std::string::npos) { i = s.find("chr"); if (i != std::string::npos) s = s.substr(i + 3); }
    else s = s.substr(i + 5);
    return s;
}

static std::vector<std::string> split_getline(const std::string& s, char d) {   /* split.cpp:3-15 */
    std::vector<std::string> v; std::stringstream ss(s); std::string item;
    while (std::getline(ss, item, d)) v.push_back(item);
    return v;
}

static long rand_int(long a, long b) { return a + (b - a) * (rand() / (RAND_MAX + 1.0)); }   /* MyDefine.cpp:290-292 */

bool SimuVars::load_fasta(const std::string& path) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) { err = "could not open " + path; return false; }
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line[0] == '>') {
            std::string name = line.substr(1);
            name = name.substr(0, name.find_first_of(" \t"));
            chroms.push_back(strip_chr(name)); seqs.push_back("");
        } else if (!seqs.empty()) seqs.back() += line;
    }
    if (chroms.empty()) { err = "ERROR: reference sequence cannot be empty!"; return false; }
    return true;
}

bool SimuVars::load_vars(const std::string& path) {   /* Genome.cpp:35-165 */
    if (path.empty()) return true;
    std::ifstream f(path.c_str());
    if (!f.is_open()) { err = "can not open file " + path; return false; }
    std::string line; int ln = 0;
    auto bad = [&](const std::string& what) { err = "ERROR: " + what + " at line " + std::to_string(ln) + " in file " + path; return false; };
    auto zyg = [&](const std::string& t, bool& het) { if (t != "homo" && t != "het") return false; het = (t == "het"); return true; };
    while (std::getline(f, line)) {
        ln++;
        if (line.empty() || line[0] == '#') continue;
        std::vector<std::string> fl = split_getline(line, '\t');
        if (fl.empty()) return bad("unrecognized aberraton type");
        const std::string& t = fl[0];
        if (t == "c") {
            if (fl.size() != 6) return bad("wrong number of fields");
            float cn = atof(fl[4].c_str()), mcn = atof(fl[5].c_str());
            if (cn < mcn) return bad("total copy number should be not lower than major copy number");
            if (cn - mcn > mcn) mcn = cn - mcn;
            cnvs[strip_chr(fl[1])].push_back({atol(fl[2].c_str()), atol(fl[3].c_str()), cn, mcn}); nCnv++;
        } else if (t == "s") {
            if (fl.size() != 6) return bad("wrong number of fields");
            if (fl[3].empty() || fl[4].empty()) return bad("empty allele");
            bool het; if (fl[3][0] == fl[4][0]) return bad("the mutated allele should be not same as the reference allele");
            if (!zyg(fl[5], het)) return bad("unrecognized SNV type");
            snvs[strip_chr(fl[1])].push_back({atol(fl[2].c_str()), fl[4][0], het}); nSnv++;
        } else if (t == "i") {
            if (fl.size() != 5) return bad("wrong number of fields");
            bool het; if (!zyg(fl[4], het)) return bad("unrecognized insert type");
            inss[strip_chr(fl[1])].push_back({atol(fl[2].c_str()), fl[3], het}); nIns++;
        } else if (t == "d") {
            if (fl.size() != 5) return bad("wrong number of fields");
            bool het; if (!zyg(fl[4], het)) return bad("unrecognized deletion type");
            dels[strip_chr(fl[1])].push_back({atol(fl[2].c_str()), atoi(fl[3].c_str()), het}); nDel++;
        } else return bad("unrecognized aberraton type");
    }
    return true;
}

static char snp_complement(char c) {   /* snp.cpp:96-110 */
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c';
        default: return 'N';
    }
}

bool SimuVars::load_snps(const std::string& path) {   /* snp.cpp:147-203 and the SNP constructor :13-36 */
    if (path.empty()) return true;
    FILE* f = fopen(path.c_str(), "r");
    if (!f) { err = "can not open SNP file " + path; return false; }
    char buf[1000];
    while (fgets(buf, 1000, f)) {
        std::vector<char*> el; el.push_back(buf);
        for (char* p = buf; *p; p++) if (*p == '\t') { *p = 0; el.push_back(p + 1); }
        if (el.size() != 6) continue;   /* the reference warns and goes on */
        std::vector<std::string> obs = split_getline(el[3], '/');
        if (obs.size() < 2 || obs[0].empty() || obs[1].empty()) continue;   /* reference: undefined behaviour */
        char strand = *el[4], ref = *el[5];
        if (strand == '-') ref = snp_complement(ref);
        char nuc = (obs[0][0] == ref) ? obs[1][0] : obs[0][0];
        if (strand == '-') nuc = snp_complement(nuc);
        snps[strip_chr(el[1])].push_back({atoll(el[2]), nuc}); nSnp++;
    }
    fclose(f);
    return true;
}

/* Genome::generateSegment, Genome.cpp:388-691. `hap[j] += edited copies of chrSeq[s-1, e)`. */
void SimuVars::segment(std::vector<std::string>& hap, const std::string& chr, const std::string& chrSeq, long s, long e, int CN, int mCN) {
    if (CN == 0) return;
    if (s - 1 < 0 || e - s + 1 < 1) throw std::runtime_error("Error: cannot construct subsequence with negative offset or length < 1");
    if (e > (long)chrSeq.size()) throw std::runtime_error("segment past the end of the chromosome (the reference reads the file bytes that follow)");
    std::string refSeq = chrSeq.substr((int)(s - 1), (int)(e - s + 1));
    for (char& c : refSeq) c = (char)toupper((unsigned char)c);
    refSeq = std::string(refSeq.c_str());   /* the reference goes through a C string */
    unsigned int refSize = (unsigned int)refSeq.size();
    if (refSize == 0) throw std::runtime_error("empty segment");
    int i, j, k, n;
    std::vector<int> mIndx, seqReps;
    auto has = [](const std::vector<int>& v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); };

    if (CN < ploidy) {   /* :411-425 — CN distinct haplotypes survive, the first mCN of them are "major" */
        for (i = 0; i < CN; i++) for (;;) { j = (int)rand_int(0, ploidy); if (!has(seqReps, j)) { seqReps.push_back(j); break; } }
        for (i = 0; i < mCN; i++) mIndx.push_back(seqReps[i]);
    } else {             /* :426-467 — every haplotype once, the CN-ploidy extra copies go to random haplotypes */
        seqReps.assign(ploidy, 1);
        n = CN - ploidy;
        k = (int)rand_int(0, ploidy);
        for (i = n; i >= 0; i--) {
            if (seqReps[k] + i == mCN) { seqReps[k] += i; mIndx.push_back(k); break; }
            else if (seqReps[k] + i == CN - mCN) { seqReps[k] += i; for (j = 0; j < ploidy; j++) if (j != k) mIndx.push_back(j); break; }
        }
        if (i >= 0 && n - i > 0 && ploidy < 2) throw std::runtime_error("endless loop in the reference (extra copies, one haplotype)");
        if (i >= 0) { n -= i; while (n > 0) { j = (int)rand_int(0, ploidy); if (j != k) { seqReps[j]++; n--; } } }
        else { while (n > 0) { j = (int)rand_int(0, ploidy); seqReps[j]++; n--; } for (i = 0; i < ploidy; i++) mIndx.push_back(i); }
    }
    std::vector<std::string> seg(ploidy);
    if (CN < ploidy) { for (i = 0; i < ploidy; i++) if (has(seqReps, i)) seg[i] = refSeq; }
    else for (i = 0; i < ploidy; i++) for (j = 0; j < seqReps[i]; j++) seg[i] += refSeq;

    /* het variants alternate between the major set (k == 0) and its complement (k == 1) */
    auto skip = [&](int kk, int jj) { bool in = has(mIndx, jj); return (kk == 0 && !in) || (kk == 1 && in); };

    k = 0;   /* SNPs :489-508 */
    auto itp = snps.find(chr);
    if (itp != snps.end()) for (const SvSnp& v : itp->second) {
        long pos = (long)v.pos;
        if (pos < s || pos > e) continue;
        int sindx = pos - s;
        for (j = 0; j < ploidy; j++) {
            if (skip(k, j)) continue;
            std::string& q = seg[j]; unsigned int len = q.length();
            for (int t = 0; t < len / refSize; t++) q[sindx + t * refSize] = v.nuc;
        }
        k = (k + 1) % 2;
    }
    k = 0;   /* SNVs :510-544 */
    auto itv = snvs.find(chr);
    if (itv != snvs.end()) for (const SvSnv& v : itv->second) {
        if (v.pos < s || v.pos > e) continue;
        int sindx = v.pos - s;
        for (j = 0; j < ploidy; j++) {
            if (v.het && skip(k, j)) continue;
            std::string& q = seg[j]; unsigned int len = q.length();
            for (int t = 0; t < len / refSize; t++) q[sindx + t * refSize] = v.alt;
        }
        if (v.het) k = (k + 1) % 2;
    }
    /* insertions :546-606 */
    std::map<int, std::map<int, int>> insAt, delAt;
    std::vector<int> insLen(ploidy, 0), delLen(ploidy, 0);
    auto shift = [&](int jj, int sindx, bool withDels) {
        int off = 0;
        for (auto& p : insAt[jj]) if (p.first <= sindx) off += p.second;
        if (withDels) for (auto& p : delAt[jj]) if (p.first <= sindx) off -= p.second;
        return off;
    };
    k = 0;
    auto iti = inss.find(chr);
    if (iti != inss.end()) for (const SvIns& v : iti->second) {
        if (v.pos < s || v.pos > e) continue;
        int sindx = v.pos - s;
        for (j = 0; j < ploidy; j++) {
            if (v.het && skip(k, j)) continue;
            int offset = shift(j, sindx, false);
            std::string& q = seg[j];
            if (refSize + insLen[j] == 0) throw std::runtime_error("division by zero in the reference");
            n = q.length() / (refSize + insLen[j]);
            int len = v.seq.length();
            for (int t = 0; t < n; t++) q.insert(sindx + offset + t * (refSize + insLen[j] + len), v.seq);
            insLen[j] += v.seq.length();
            insAt[j].insert(std::make_pair(sindx, (int)v.seq.length()));
        }
        if (v.het) k = (k + 1) % 2;
    }
    /* deletions :608-679 — k is NOT reset here (it carries over from the insertion loop) */
    auto itd = dels.find(chr);
    if (itd != dels.end()) for (const SvDel& v : itd->second) {
        if (v.pos < s || v.pos > e) continue;
        int sindx = v.pos - s;
        int dl = v.len;
        for (j = 0; j < ploidy; j++) {
            if (v.het && skip(k, j)) continue;
            int offset = shift(j, sindx, true);
            if (sindx + offset < 0) continue;
            std::string& q = seg[j];
            if (refSize + insLen[j] - delLen[j] == 0) throw std::runtime_error("division by zero in the reference");
            n = q.length() / (refSize + insLen[j] - delLen[j]);
            for (int t = 0; t < n; t++) q.erase(sindx + offset + t * (refSize + insLen[j] - delLen[j] - dl), dl);
            delLen[j] += dl;
            delAt[j].insert(std::make_pair(sindx, dl));
        }
        if (v.het) k = (k + 1) % 2;
    }
    for (i = 0; i < ploidy; i++) {
        for (char& c : seg[i]) c = (char)toupper((unsigned char)c);
        hap[i] += seg[i];
    }
}

bool SimuVars::run(std::string& out) {   /* Genome::saveSequence, Genome.cpp:329-386 */
    srand(seed);
    int mCN = (int)ceil((float)ploidy / 2);
    out.clear();
    try {
        for (size_t c = 0; c < chroms.size(); c++) {
            const std::string& chr = chroms[c];
            const long chrLen = (long)seqs[c].size();
            std::vector<std::string> hap(ploidy);
            long segStart = 1;
            auto itc = cnvs.find(chr);
            if (itc != cnvs.end()) for (SvCnv cnv : itc->second) {
                if (segStart > chrLen) break;
                cnv.epos = std::min(cnv.epos, chrLen);
                if (segStart < cnv.spos) segment(hap, chr, seqs[c], segStart, cnv.spos - 1, ploidy, mCN);
                segment(hap, chr, seqs[c], cnv.spos, cnv.epos, (int)cnv.cn, (int)cnv.mcn);
                segStart = cnv.epos + 1;
            }
            if (segStart <= chrLen) segment(hap, chr, seqs[c], segStart, chrLen, ploidy, mCN);
            for (int j = 0; j < ploidy; j++) {
                out += ">" + chr + "_" + std::to_string(j + 1) + "_" + std::to_string(chrLen) + "\n";
                unsigned int sindx = 0, length = hap[j].length();
                while (sindx < length) { out += hap[j].substr(sindx, 100); out += "\n"; sindx += 100; }
            }
        }
    } catch (const std::exception& ex) {   /* the reference terminates on these (uncaught std::out_of_range / exit(1)) */
        err = std::string("reference aborts: ") + ex.what();
        return false;
    }
    return true;
}

}  // namespace orc
