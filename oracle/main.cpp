/* TEST INFRASTRUCTURE — CPU oracle driver (see draws.h header).
 *
 *   scs_oracle genreads -i <fa> -m <profile> [-p N] [-r gamma] [-l PE|SE] [-c cov] [-s isize]
 *                       -o <prefix> (--tape <log prefix> | --seed <u64>) [--dump <prefix>] [--no-fastq]
 *
 * --tape replays the draw logs of oracle/_ref/bin/scssim_replay (-t 1); the FASTQ must then be
 * byte-identical to the reference's. --dump writes every intermediate array (and, in tape mode,
 * the per-entity tape offsets) as raw little-endian binaries for the GPU parity tests.
 * Also exported as `orc_main` from liboracle.so for in-process use.
 */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>

#include "genreads.h"
#include "simuvars.h"

using namespace orc;

template <class T> static void dump_vec(const std::string& path, const std::vector<T>& v) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "oracle: cannot write %s\n", path.c_str()); exit(2); }
    if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}
static void dump_str(const std::string& path, const std::string& s) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "oracle: cannot write %s\n", path.c_str()); exit(2); }
    if (!s.empty()) fwrite(s.data(), 1, s.size(), f);
    fclose(f);
}

/*   scs_oracle simuvars -r <ref.fa> [-s <snp file>] [-v <variation file>] -o <out.fa> [--seed <libc seed, default 1>] */
static int simuvars_main(int argc, char** argv) {
    SimuVars sv; std::string ref, snp, var, out;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "-r" || a == "--ref") ref = val();
        else if (a == "-s" || a == "--snp") snp = val();
        else if (a == "-v" || a == "--var") var = val();
        else if (a == "-o" || a == "--output") out = val();
        else if (a == "--seed") sv.seed = (unsigned)strtoul(val(), NULL, 0);
        else if (a == "--ploidy") sv.ploidy = atoi(val());
        else { fprintf(stderr, "oracle: unknown option %s\n", a.c_str()); return 1; }
    }
    if (ref.empty() || out.empty()) { fprintf(stderr, "oracle: missing arguments\n"); return 1; }
    auto t0 = std::chrono::steady_clock::now();
    std::string text;
    if (!sv.load_vars(var) || !sv.load_snps(snp) || !sv.load_fasta(ref) || !sv.run(text)) { fprintf(stderr, "oracle: %s\n", sv.err.c_str()); return 3; }
    auto t1 = std::chrono::steady_clock::now();
    dump_str(out, text);
    fprintf(stderr, "{\"cnv\": %ld, \"snv\": %ld, \"ins\": %ld, \"del\": %ld, \"snp\": %ld, \"bytes\": %zu, \"t_run\": %.4f}\n", sv.nCnv, sv.nSnv, sv.nIns,
            sv.nDel, sv.nSnp, text.size(), std::chrono::duration<double>(t1 - t0).count());
    return 0;
}

extern "C" int orc_main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "simuvars") == 0) return simuvars_main(argc, argv);
    if (argc < 2 || strcmp(argv[1], "genreads") != 0) { fprintf(stderr, "usage: scs_oracle genreads|simuvars ...\n"); return 1; }
    Sim sim; std::string fa, model, out, tape, dump; uint64_t seed = 0; bool haveSeed = false, noFastq = false;
    for (int i = 2; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "-i" || a == "--input") fa = val();
        else if (a == "-m" || a == "--model") model = val();
        else if (a == "-o" || a == "--output") out = val();
        else if (a == "-p" || a == "--primers") sim.P.primers = atol(val());
        else if (a == "-r" || a == "--gamma") sim.P.gamma = atof(val());
        else if (a == "-c" || a == "--coverage") sim.P.coverage = atof(val());
        else if (a == "-s" || a == "--isize") sim.P.isize = atoi(val());
        else if (a == "-l" || a == "--layout") sim.P.paired = std::string(val()) != "SE";
        else if (a == "-t" || a == "--threads") val();
        else if (a == "--tape") tape = val();
        else if (a == "--seed") { seed = strtoull(val(), NULL, 0); haveSeed = true; }
        else if (a == "--dump") dump = val();
        else if (a == "--no-fastq") noFastq = true;
        else { fprintf(stderr, "oracle: unknown option %s\n", a.c_str()); return 1; }
    }
    if (fa.empty() || model.empty() || (out.empty() && !noFastq) || (tape.empty() && !haveSeed)) { fprintf(stderr, "oracle: missing arguments\n"); return 1; }
    if (!sim.load_fasta(fa)) { fprintf(stderr, "oracle: %s\n", sim.err.c_str()); return 1; }
    if (!sim.prof.load(model, sim.P.paired, sim.P.isize)) { fprintf(stderr, "oracle: %s\n", sim.prof.err.c_str()); return 1; }
    Draws* d = tape.empty() ? (Draws*)new PhiloxDraws(seed) : (Draws*)new TapeDraws(tape);
    sim.D = d;
    auto t0 = std::chrono::steady_clock::now();
    sim.split_to_frags();
    auto t1 = std::chrono::steady_clock::now();
    sim.amplify();
    auto t2 = std::chrono::steady_clock::now();
    sim.yield_reads();
    auto t3 = std::chrono::steady_clock::now();
    if (!sim.err.empty()) { fprintf(stderr, "%s\n", sim.err.c_str()); return 1; }
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    bool consumed = true;
    if (!tape.empty()) consumed = ((TapeDraws*)d)->fully_consumed();
    fprintf(stderr, "{\"frags\": %zu, \"semis\": %zu, \"fulls\": %zu, \"reads\": %llu, \"records\": %llu, \"t_frag\": %.4f, \"t_amplify\": %.4f, \"t_reads\": %.4f, \"tapes_consumed\": %s}\n",
            sim.frags.size(), sim.semis.size(), sim.fulls.size(), (unsigned long long)sim.reads, (unsigned long long)sim.nRecords,
            secs(t0, t1), secs(t1, t2), secs(t2, t3), consumed ? "true" : "false");
    if (!noFastq) {
        if (sim.P.paired) { dump_str(out + "_1.fq", sim.fq1); dump_str(out + "_2.fq", sim.fq2); }
        else dump_str(out + ".fq", sim.fq1);
    }
    if (!dump.empty()) {
        std::vector<int64_t> fr; for (auto& f : sim.frags) { fr.push_back(f.seq); fr.push_back(f.start0); fr.push_back(f.len); fr.push_back(f.strand); fr.push_back(f.primers); }
        dump_vec(dump + ".frags.i64", fr);
        auto amps = [&](const std::vector<Amp>& v) { std::vector<uint32_t> o; for (auto& a : v) { o.push_back(a.tmpl); o.push_back(a.spos); o.push_back(a.len); o.push_back(a.gc); o.push_back(a.primers); o.push_back(a.err_off); o.push_back(a.err_n); } return o; };
        dump_vec(dump + ".semis.u32", amps(sim.semis));
        dump_vec(dump + ".fulls.u32", amps(sim.fulls));
        std::vector<uint32_t> er; for (auto& e : sim.errs) { er.push_back(e.pos); er.push_back(e.alt); }
        dump_vec(dump + ".errs.u32", er);
        dump_vec(dump + ".counts.u32", sim.readNumbers);
        dump_vec(dump + ".weights.f64", sim.weights);
        dump_vec(dump + ".gcf.f64", sim.gcFactors);
        dump_vec(dump + ".semi_batches.u64", sim.semiBatchEnd);
        dump_vec(dump + ".full_batches.u64", sim.fullBatchEnd);
        std::vector<int64_t> pc(sim.primerCount.begin(), sim.primerCount.end());
        dump_vec(dump + ".primer_counts.i64", pc);
        std::vector<uint64_t> meta = {sim.reads, sim.totalPrimers, sim.nRecords};
        dump_vec(dump + ".meta.u64", meta);
        if (!tape.empty()) {
            TapeDraws* td = (TapeDraws*)d;
            for (int dm = 0; dm < D_COUNT; dm++) {
                std::vector<uint64_t> m; for (auto& k : td->marks[dm]) { m.push_back(k.entity); m.push_back(k.off_real); m.push_back(k.off_int); }
                dump_vec(dump + ".marks" + std::to_string(dm) + ".u64", m);
            }
        }
    }
    delete d;
    return consumed ? 0 : 4;
}

#ifndef ORC_NO_MAIN
int main(int argc, char** argv) { return orc_main(argc, argv); }
#endif
