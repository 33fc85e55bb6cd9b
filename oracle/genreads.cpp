/* TEST INFRASTRUCTURE — CPU oracle; see genreads.h for the reference map. */
#include "genreads.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <cstdlib>

namespace orc {

static const char* BASES = "ACGT";

/* ------------------------------------------------------------------ FASTA */
/* Names as the reference's .fai reader keys them (lib/fastahack/Fasta.cpp:57-68): first
 * token of the header, with everything up to and including "chrom"/"chr" removed. */
static std::string ref_seq_name(const std::string& header) {
    std::string name = header.substr(0, header.find_first_of(" \t"));
    size_t i = name.find("chrom");
    if (i == std::string::npos) { i = name.find("chr"); if (i != std::string::npos) name = name.substr(i + 3); }
    else name = name.substr(i + 5);
    return name;
}

bool Sim::load_fasta(const std::string& path) {
    std::ifstream f(path.c_str());
    if (!f.is_open()) { err = "could not open " + path; return false; }
    std::string line;
    while (std::getline(f, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (line.empty()) continue;
        if (line[0] == '>') { names.push_back(ref_seq_name(line.substr(1))); seqs.push_back(""); }
        else if (!seqs.empty()) {
            for (char& c : line) c = (char)toupper((unsigned char)c);   /* Genome::getSubSequence, Genome.cpp:274 */
            seqs.back() += line;
        }
    }
    if (seqs.empty()) { err = "reference sequence cannot be empty"; return false; }
    return true;
}

/* -------------------------------------------------------------- fragments */
void Sim::split_to_frags() {   /* Genome.cpp:753-782 */
    frags.clear();
    for (size_t c = 0; c < seqs.size(); c++) {
        long chrLen = (long)seqs[c].size();
        long start = 1;
        D->begin(D_FRAG, c);
        while (start <= chrLen) {
            int fragLen = (int)uni_int(D->next(E_REAL), P.fragMin, P.fragMax + 1);
            if (start + fragLen - 1 > chrLen) break;
            frags.push_back({(int)c, start - 1, fragLen, 1, 0});
            frags.push_back({(int)c, start - 1, fragLen, -1, 0});
            start += fragLen;
        }
        if (start <= chrLen) {   /* tail: both copies strand +1 (quirk) */
            int len = (int)(chrLen - start + 1);
            frags.push_back({(int)c, start - 1, len, 1, 0});
            frags.push_back({(int)c, start - 1, len, 1, 0});
        }
    }
}

/* Fragment::createSequence stores reverse(G) for strand +1 and complement(G) for strand -1
 * (Fragment.cpp:42-47); amplification works on complement(stored) (Fragment.cpp:65-68). */
char Sim::frag_tmpl(const Frag& f, uint32_t i) const {
    const std::string& g = seqs[f.seq];
    if (f.strand == 1) return complement_base(g[f.start0 + f.len - 1 - i]);
    return complement_base(complement_base(g[f.start0 + i]));
}

void Sim::semi_window(const Amp& s, std::string& out) const {   /* Amplicon.cpp:341-376, before the reverse */
    const Frag& f = frags[s.tmpl];
    out.resize(s.len);
    for (uint32_t j = 0; j < s.len; j++) out[j] = frag_tmpl(f, s.spos + j);
    for (uint32_t e = 0; e < s.err_n; e++) out[errs[s.err_off + e].pos] = BASES[errs[s.err_off + e].alt];
}

void Sim::semi_tmpl(const Amp& s, std::string& out) const {    /* Amplicon.cpp:170-171: complement(getSequence()) */
    std::string w; semi_window(s, w);
    out.resize(s.len);
    for (uint32_t i = 0; i < s.len; i++) out[i] = complement_base(w[s.len - 1 - i]);
}

void Sim::full_sequence(const Amp& a, std::string& out) const {   /* Amplicon.cpp:266-340 */
    std::string u; semi_tmpl(semis[a.tmpl], u);
    for (uint32_t e = 0; e < a.err_n; e++) u[a.spos + errs[a.err_off + e].pos] = BASES[errs[a.err_off + e].alt];
    out = u.substr(a.spos, a.len);
}

/* ---------------------------------------------------------------- primers */
void Sim::create_primers() {   /* Malbac.cpp:36-81 */
    primerCount.assign(65536, P.primers);
    totalPrimers = (uint64_t)65536 * (uint64_t)P.primers;
}

int Sim::take_primer(const char* s8) {   /* updatePrimerCount(s,-1), Malbac.cpp:91-103 */
    int idx = 0;
    for (int i = 0; i < 8; i++) {
        int b = base_index(s8[i]);
        if (b < 0) return getenv("ORC_N_BINDS") ? 1 : 0;   /* reference: uninitialised counter (UB); here: no primer binds an N site */
        idx = idx * 4 + b;
    }
    if (primerCount[idx] - 1 < 0) return 0;
    primerCount[idx] -= 1;
    return 1;
}

long Sim::poiss_rand(double lambda) {   /* MyDefine.cpp:69-80 */
    long x = -1; double log1 = 0, log2 = -lambda;
    do {
        double u = uni_real(D->next(E_REAL), 0, 1);
        log1 += D->is_tape() ? log(u) : det_log(u);
        x++;
    } while (log1 >= log2);
    return x;
}

void Sim::set_primers(bool onlyFrags, int round) {   /* Malbac.cpp:236-283 */
    uint64_t templateNum = frags.size(); double totalLen = 0;
    for (auto& f : frags) totalLen += (unsigned)f.len;
    if (!onlyFrags) { templateNum += semis.size(); for (auto& s : semis) totalLen += s.len; }
    uint64_t expectedPrimers = (uint64_t)((double)totalPrimers * P.gamma * (double)templateNum);
    uint64_t count = 0, t = 0;
    for (auto& f : frags) {
        double lambda = (double)expectedPrimers * (1.0 * (unsigned)f.len / totalLen);
        D->begin(D_POIS, ((uint64_t)round << 40) | t++);
        long k = poiss_rand(lambda);
        count += k; f.primers = (int)k;
    }
    if (!onlyFrags) for (auto& s : semis) {
        double lambda = (double)expectedPrimers * (1.0 * s.len / totalLen);
        D->begin(D_POIS, ((uint64_t)round << 40) | t++);
        long k = poiss_rand(lambda);
        count += k; s.primers = (uint16_t)(k & 0xFFF);   /* 12-bit field, Amplicon.cpp:76-79 */
    }
    totalPrimers -= count;
}

/* ---------------------------------------------------------- amplification */
/* Shared body of Fragment::amplify (Fragment.cpp:52-137) and Amplicon::amplify (Amplicon.cpp:156-240). */
template <class GetBase>
void Sim::amplify_template(GetBase tb, uint32_t length, int primerNum, bool fromFragment, uint32_t tmplIdx,
                           std::vector<Amp>& out) {
    if ((int)length < P.ampMin + 27) return;
    std::vector<uint8_t> attached(length, 0);
    for (int i = 0; i < primerNum; i++) {
        int tries = 0; uint32_t spos = 0, alen = 0;
        for (;;) {
            spos = (uint32_t)uni_int(D->next(E_INT), 27, length);
            alen = (uint32_t)uni_real(D->next(E_REAL), P.ampMin, P.ampMax + 1);
            tries++;
            if (tries > 50) break;
            if (spos + alen > length || attached[spos]) continue;
            char s8[8]; for (int k = 0; k < 8; k++) s8[k] = tb(spos + k);
            if (take_primer(s8) == 1) break;
        }
        if (tries > 50) break;   /* abandon the template's remaining primers */
        attached[spos] = 1;
        int gcNum = 0; bool anyN = false;   /* countGC, MyDefine.cpp:434-452 */
        for (uint32_t j = 0; j < alen; j++) { char c = tb(spos + j); if (c == 'G' || c == 'C') gcNum++; else if (c == 'N') anyN = true; }
        if (anyN) gcNum = 0;
        uint32_t eoff = (uint32_t)errs.size(), en = 0;
        auto substitute = [&](uint32_t j) {   /* Fragment.cpp:108-122 / Amplicon.cpp:211-225 */
            char base = tb(spos + j); unsigned n;
            do {
                n = fromFragment ? (unsigned)uni_int(D->next(E_INT), 0, 4)        /* Fragment.cpp:110 */
                                 : (unsigned)uni_real(D->next(E_REAL), 0, 4);    /* Amplicon.cpp:213 */
            } while (BASES[n] == base);
            if (BASES[n] == 'C' || BASES[n] == 'G') gcNum++;
            if (base == 'C' || base == 'G') gcNum--;
            errs.push_back({j, (uint8_t)n}); en++;
        };
        if (D->is_tape()) {
            /* the reference's own loop: one draw per base, j = 8 .. alen-1 (Fragment.cpp:105-107) */
            for (uint32_t j = 8; j < alen; j++) {
                double p = uni_real(D->next(E_REAL), 0, 1);
                if (p < P.ber) substitute(j);
            }
        } else {
            /* free-running streams: the distance to the next error is drawn directly, P(gap = g) = (1-ber)^g * ber — the same
             * distribution as one Bernoulli(ber) draw per base, at one draw per error instead of one per base. The CUDA path
             * (amplify.cu) defines its Philox streams the same way; det_log keeps host and device bit-identical. */
            const double l1p = det_log(1.0 - P.ber);
            uint64_t j = 7;
            for (;;) {
                double u = ((double)D->next(E_REAL) + 0.5) / 4294967296.0;
                double g = floor(det_log(u) / l1p);
                if (!(g < (double)alen)) break;
                j += 1 + (uint64_t)g;
                if (j >= alen) break;
                substitute((uint32_t)j);
            }
        }
        out.push_back({tmplIdx, spos, alen, (uint32_t)std::max(0, gcNum), 0, eoff, en});
    }
}

void Sim::amplify_frags(int pass) {   /* Malbac.cpp:318-343 with one task (-t 1) */
    std::vector<Amp> batch;
    for (size_t i = 0; i < frags.size(); i++) {
        const Frag& f = frags[i];
        D->begin(D_AMPF, ((uint64_t)pass << 40) | i);
        amplify_template([&](uint32_t k) { return frag_tmpl(f, k); }, (uint32_t)f.len, f.primers, true, (uint32_t)i, batch);
    }
    /* insertLinkList prepends (Amplicon.cpp:574-585); extendSemiAmplicons appends the task list */
    semis.insert(semis.end(), batch.rbegin(), batch.rend());
    semiBatchEnd.push_back(semis.size());
}

void Sim::amplify_semis(int cycle) {   /* Malbac.cpp:345-368 with one task */
    std::vector<Amp> batch; std::string u;
    size_t n = semis.size();
    for (size_t i = 0; i < n; i++) {
        D->begin(D_AMPS, ((uint64_t)cycle << 40) | i);
        if ((int)semis[i].len < P.ampMin + 27) continue;
        semi_tmpl(semis[i], u);
        amplify_template([&](uint32_t k) { return u[k]; }, semis[i].len, semis[i].primers, false, (uint32_t)i, batch);
    }
    fulls.insert(fulls.end(), batch.rbegin(), batch.rend());
    fullBatchEnd.push_back(fulls.size());
}

void Sim::amplify() {   /* Malbac.cpp:173-201 */
    create_primers();
    set_primers(true, 0);
    amplify_frags(0);
    const bool dbg = getenv("ORC_DEBUG") != nullptr;
    if (dbg) fprintf(stderr, "[orc] round 0: semis %zu\n", semis.size());
    for (int i = 0; i < 5; i++) {
        if (totalPrimers == 0) break;
        set_primers(false, i + 1);
        amplify_semis(i + 1);
        if (i < 4) amplify_frags(i + 1);
        if (dbg) fprintf(stderr, "[orc] cycle %d: semis %zu fulls %zu\n", i + 1, semis.size(), fulls.size());
    }
}

/* ------------------------------------------------------------ read counts */
void Sim::set_read_counts() {   /* Malbac.cpp:370-408 */
    size_t n = fulls.size();
    weights.assign(n, 0); gcFactors.assign(n, 0); readNumbers.assign(n, 0);
    if (n == 0) return;
    for (size_t i = 0; i < n; i++) {
        int gc = (int)(100u * fulls[i].gc / fulls[i].len);   /* Amplicon.cpp:396-400 */
        D->begin(D_GCF, i);
        double f = (gc < 0 || gc > 100) ? 0.0 : D->gc_factor(prof.gcMeans[gc], prof.gcStd);
        gcFactors[i] = f;
        weights[i] = f * fulls[i].len / (double)((unsigned)P.fragSize * (unsigned)P.fragSize);
    }
    double S = 0; for (size_t i = 0; i < n; i++) S += weights[i];
    for (size_t i = 0; i < n; i++) weights[i] /= (ZERO_FINAL + S);
    long r = (long)reads; unsigned long sum = 0;
    for (size_t i = 0; i < n; i++) { unsigned rc = (unsigned)(weights[i] * r); readNumbers[i] = rc; sum += rc; }
    r -= (long)sum;
    /* randIndx_hp(wls, reads, readNumbers, true), MyDefine.cpp:203-272, threads = 1 */
    unsigned long rem = (unsigned long)r;
    unsigned ac = (unsigned)n;
    unsigned load = std::max(1u, std::min(1000u, ac));
    struct Chunk { unsigned s, m; double total; double samples; };
    std::vector<Chunk> chunks; unsigned long count = 0;
    for (unsigned s = 0; s < ac; s += load) {
        unsigned e = (s + load > ac) ? ac - 1 : s + load - 1;
        double total = 0; for (unsigned i = s; i <= e; i++) total += weights[i];
        double ns = (unsigned)(total * rem);
        count += (unsigned long)ns;
        chunks.push_back({s, e - s + 1, total, ns});
    }
    rem -= count;
    if (rem > 0) {
        std::vector<double> probs(chunks.size());
        probs[0] = chunks[0].total; for (size_t i = 1; i < chunks.size(); i++) probs[i] = probs[i - 1] + chunks[i].total;
        D->begin(D_MULTM, 0);
        while (rem-- > 0) chunks[Profile::rand_index(probs.data(), (unsigned)probs.size(), D->next(E_REAL))].samples += 1;
    }
    std::vector<double> cdf;
    for (size_t c = 0; c < chunks.size(); c++) {   /* batchSampling, MyDefine.cpp:191-201 */
        const Chunk& ch = chunks[c];
        cdf.assign(ch.m, 0); double prev = 0;
        for (unsigned k = 0; k < ch.m; k++) { cdf[k] = prev + weights[ch.s + k] / ch.total; prev = cdf[k]; }
        D->begin(D_MULTC, c);
        unsigned ns = (unsigned)ch.samples;
        for (unsigned i = 0; i < ns; i++) readNumbers[ch.s + Profile::rand_index(cdf.data(), ch.m, D->next(E_REAL))] += 1;
    }
    if (P.paired) { int k = 1; for (size_t i = 0; i < n; i++) if (readNumbers[i] % 2 == 1) { readNumbers[i] += k; k *= -1; } }
}

/* ------------------------------------------------------------------ reads */
void Sim::append_record(std::string& out, long ampIdx, int fragCount, const char* suffix, const std::string& s,
                        const std::string& q) {   /* Amplicon.cpp:459-468, 497-505 */
    char head[64];
    snprintf(head, sizeof head, "@%d#%d%s\n", (int)ampIdx, fragCount, suffix);
    out += head; out += s; out += "\n+\n"; out += q; out += "\n";
}

void Sim::yield_reads() {   /* Malbac.cpp:410-460, Amplicon.cpp:402-565 */
    unsigned long refLen = 0;
    for (auto& nm : names) { auto fs = split(nm, '_'); if (!fs.empty()) refLen += atoi(fs.back().c_str()); }
    refLen /= 2;
    reads = (unsigned long)(refLen * P.coverage / prof.readLength);
    set_read_counts();
    const int RL = prof.readLength;
    std::vector<uint64_t> slotBase(fulls.size() + 1, 0);
    for (size_t a = 0; a < fulls.size(); a++) slotBase[a + 1] = slotBase[a] + (P.paired ? readNumbers[a] / 2 : readNumbers[a]);
    std::string amp, r1s, r1q, r2s, r2q, mate;
    fq1.clear(); fq2.clear(); nRecords = 0;
    for (size_t a = 0; a < fulls.size(); a++) {
        int n = (int)readNumbers[a];
        if (n == 0) continue;
        full_sequence(fulls[a], amp);
        int ampLen = (int)amp.size();
        if (ampLen < RL) continue;
        int fragCount = 0, failCount = 0; uint64_t slot = 0; bool fresh = true;
        while (n > 0) {
            fragCount++;
            if (fresh) { D->begin(D_READ, slotBase[a] + slot); fresh = false; }
            if (!P.paired) {
                long pos = uni_int(D->next(E_INT), 0, ampLen - RL + 1);
                prof.predict(amp.substr(pos, RL), true, *D, r1s, r1q);
                append_record(fq1, (long)a, fragCount, "", r1s, r1q);
                n--; nRecords++;
            } else {
                if (!prof.hasISize) { err = "Error: unrecognized parameter name \"insertSize\""; return; }   /* Profile.cpp:1484 */
                int isz = prof.yield_insert_size(*D);
                if (isz < RL || isz > ampLen) { failCount++; if (failCount > 1000) break; continue; }
                long pos = uni_int(D->next(E_INT), 0, ampLen - isz + 1);
                prof.predict(amp.substr(pos, RL), true, *D, r1s, r1q);
                mate.resize(RL);
                for (int i = 0; i < RL; i++) mate[i] = complement_base(amp[pos + isz - 1 - i]);
                prof.predict(mate, false, *D, r2s, r2q);
                append_record(fq1, (long)a, fragCount, "/1", r1s, r1q);
                append_record(fq2, (long)a, fragCount, "/2", r2s, r2q);
                n -= 2; nRecords += 2;
            }
            slot++; fresh = true;
        }
    }
}

}  // namespace orc
