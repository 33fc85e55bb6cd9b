// `scssim genreads` and `scssim simuvars` drop-in: same flags, defaults, validation messages, stderr progress lines and
// output files as the reference's CLI (/root/reference/src/scssim.cpp:23-76,109-172,285-404,421-485), driving
// the CUDA path through the C ABI. `-t` (the reference's worker-thread count) sets the number of host
// threads that write the FASTQ slabs to the files: the compute runs on the GPU. Extra long-only flags: --seed <u64>, --device <n>.
#include <fcntl.h>
#include <getopt.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cerrno>

#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "scssim_b200.h"

using namespace std;

// .gz input is expanded beside the input like the reference does (lib/genome/Genome.cpp:183-187), but without a shell:
// gzip runs through fork/execvp with its stdout redirected, so paths with spaces or shell metacharacters are just paths.
static bool gunzip_beside(const string& gz, string* out) {
    *out = gz.substr(0, gz.size() - 3);
    int fd = open(out->c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) return false;
    pid_t pid = fork();
    if (pid < 0) { close(fd); return false; }
    if (pid == 0) {
        dup2(fd, 1); close(fd);
        execlp("gzip", "gzip", "-cd", "--", gz.c_str(), (char*)NULL);
        _exit(127);
    }
    close(fd);
    int st = 0;
    while (waitpid(pid, &st, 0) < 0) if (errno != EINTR) return false;
    return WIFEXITED(st) && WEXITSTATUS(st) == 0;
}
static bool is_gz(const string& p) { return p.size() > 3 && p.compare(p.size() - 3, 3, ".gz") == 0; }

static void usage(const char* app) {
    cerr << "\nSCSsim version: 1.0 (B200-native genreads)" << endl;
    cerr << "Usage: " << app << " [subcommand] [options]" << endl << endl
         << "Optional arguments:" << endl
         << "    -h, --help                      give this information" << endl
         << "    -v, --version <string>          print software version" << endl << endl
         << "Available subcmds:" << endl
         << "    simuvars          simulate the genome sequence of single cells" << endl
         << "    genreads          simulate sequencing reads of single cell" << endl << endl;
}

static void usage_simuVars(const char* app) {
    cerr << "Usage: scssim " << app << " [options]" << endl << endl
         << "Options:" << endl
         << "    -h, --help                      give this information" << endl
         << "    -r, --ref <string>              reference file (.fasta)" << endl
         << "    -s, --snp <string>              SNP file containing the SNPs to be simulated [Default:null]" << endl
         << "    -v, --var <string>              variation file containing the genomic variations to be simulated [Default:null]" << endl
         << "    -o, --output <string>           output file (.fasta) to save generated sequences" << endl
         << "        --device <int>              CUDA device ordinal [Default:0]" << endl << endl
         << "Example:" << endl
         << "    scssim " << app << " -r /path/to/hg19.fa -s /path/to/hg19.snp138.1based.txt -v /path/to/variation.txt -o /path/to/results.fa" << endl << endl;
}

static int die(scs_ctx* c, int rc);

// `scssim simuvars`: parseArgs_simuVars + main's simuvars branch (src/scssim.cpp:33-38,109-172)
static int main_simuvars(int argc, char* argv[], time_t start_t) {
    string refFile, snpFile, varFile, outFile; int device = 0;
    struct option long_options[] = {{"help", no_argument, 0, 'h'}, {"ref", required_argument, 0, 'r'}, {"snp", required_argument, 0, 's'},
                                    {"var", required_argument, 0, 'v'}, {"output", required_argument, 0, 'o'}, {"device", required_argument, 0, 1001},
                                    {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "hr:s:v:o:", long_options, NULL)) != -1) {
        switch (c) {
            case 'h': usage_simuVars(argv[0]); return 0;
            case 'r': refFile = optarg; break;
            case 's': snpFile = optarg; break;
            case 'v': varFile = optarg; break;
            case 'o': outFile = optarg; break;
            case 1001: device = atoi(optarg); break;
            default: usage_simuVars(argv[0]); return 1;
        }
    }
    if (refFile.empty()) { cerr << "Use --ref to specify the reference file (fasta)." << endl; usage_simuVars(argv[0]); return 1; }
    if (snpFile.empty()) cerr << "Warning: SNP file not specified!" << endl << "No SNPs will be inserted into the genome." << endl;
    if (varFile.empty()) cerr << "Warning: variation file not specified!" << endl << "No variations will be inserted into the genome." << endl;
    if (outFile.empty()) { cerr << "Use --output to specify the output file." << endl; usage_simuVars(argv[0]); return 1; }
    scs_params P; scs_default_params(&P); P.device = device;
    scs_ctx* ctx = nullptr;
    int rc = scs_create(&P, &ctx);
    if (rc) return die(nullptr, rc);
    string fa = refFile;   // .gz input: gunzipped beside the input (lib/genome/Genome.cpp:183-187)
    if (is_gz(fa)) {
        string plain;
        if (!gunzip_beside(fa, &plain)) { cerr << "could not open " << fa << endl; scs_destroy(ctx); return 1; }
        fa = plain;
    }
    scs_simuvars_params sp; scs_simuvars_default_params(&sp);
    rc = scs_simuvars(ctx, &sp, fa.c_str(), snpFile.c_str(), varFile.c_str(), outFile.c_str());
    scs_simuvars_stats st; scs_simuvars_get_stats(ctx, &st);
    cerr << scs_simuvars_warnings(ctx);
    if (rc) return die(ctx, rc);
    // what Genome::loadAbers / loadSNPs / loadRefSeq report (Genome.cpp:160-164,173,197)
    if (!varFile.empty()) cerr << "\nDetails of the aberrations loaded from file " << varFile << " are as follows:" << endl << "CNV: " << st.n_cnv << endl
                               << "SNV: " << st.n_snv << endl << "Insert: " << st.n_ins << endl << "Deletion: " << st.n_del << endl;
    if (!snpFile.empty()) cerr << "\n" << st.n_snp << " SNPs to simulate were loaded from file " << snpFile << endl;
    cerr << "\nReference sequence was loaded from file " << refFile << endl;
    scs_destroy(ctx);
    long used = (long)(time(NULL) - start_t);
    cerr << "\nElapsed time: " << used / 60 << " minutes and " << used % 60 << " seconds!\n" << endl;
    return 0;
}

static void usage_genReads(const char* app) {
    cerr << "Usage: scssim " << app << " [options]" << endl << endl
         << "Options:" << endl
         << "    -h, --help                      give this information" << endl
         << "    -i, --input <string>            sequence file (.fasta) generated by simuVars program" << endl
         << "  MALBAC options:" << endl
         << "    -p, --primers <int>             the number of primers [Default:100000]" << endl
         << "    -r, --gamma <float>             a parameter controlling the number of primers used in each cycle [Default:1e-9]" << endl
         << "  Read simulation options:" << endl
         << "    -m, --model <string>            profile inferred from real sequencing data" << endl
         << "    -l, --layout <string>           read layout (SE for single end, PE for paired-end) [Default:PE]" << endl
         << "    -c, --coverage <float>          sequencing coverage [Default:5]" << endl
         << "    -s, --isize <int>               mean insert size for paired-end sequencing [Default:260]" << endl
         << "    -t, --threads <int>             number of threads to use [Default:1] (host threads writing the files; compute runs on the GPU)" << endl
         << "    -o, --output <string>           the prefix of output file" << endl
         << "        --seed <int>                random seed [Default:time]" << endl
         << "        --device <int>              CUDA device ordinal [Default:0]" << endl
         << "        --gpus <int>                shard the cell's sequences over this many GPUs [Default:1]" << endl
         << "        --gz                        write block-gzip compressed FASTQ (<prefix>_1.fq.gz ...; with --gpus N one shard per GPU," << endl
         << "                                    <prefix>.rank<r>_1.fq.gz, whose concatenation in rank order is the output)" << endl << endl
         << "Example:" << endl
         << "    scssim " << app << " -i /path/to/ref.fa -m /path/to/hiseq2500.profile -t 5 -o /path/to/reads" << endl << endl;
}

// --gpus N: one host thread per GPU in this process, collectives over NCCL inside the library (scs_nccl_init). ThreadSum is the
// stand-in for the single-GPU test hook only (SCS_CLI_SAME_DEVICE=1): plain sums through shared memory behind a barrier.
struct ThreadSum {
    int world; std::mutex mu; std::condition_variable cv; int arrived = 0; long gen = 0; bool aborted = false;
    std::vector<uint64_t> au; std::vector<double> ad;
    explicit ThreadSum(int w) : world(w) {}
    // a worker that fails calls this before it returns: everybody waiting in (or arriving at) a sum gets a non-zero result
    void abort() { std::lock_guard<std::mutex> lk(mu); aborted = true; cv.notify_all(); }
    template <class T> int sum(std::vector<T>& acc, T* buf, size_t n) {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return 1;
        if (arrived == 0) acc.assign(n, T(0));
        for (size_t i = 0; i < n; i++) acc[i] += buf[i];
        long g = gen;
        if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g || aborted; });
        if (aborted) return 1;
        for (size_t i = 0; i < n; i++) buf[i] = acc[i];
        // second phase: nobody may start the next sum (and clear acc) before everyone has copied the result out
        long g2 = gen;
        if (++arrived == world) { arrived = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g2 || aborted; });
        return aborted ? 1 : 0;
    }
};
static int sum_u64(void* u, uint64_t* b, size_t n) { ThreadSum* t = (ThreadSum*)u; return t->sum(t->au, b, n); }
static int sum_f64(void* u, double* b, size_t n) { ThreadSum* t = (ThreadSum*)u; return t->sum(t->ad, b, n); }

static int die(scs_ctx* c, int rc) {
    cerr << scs_last_error(c) << endl;
    if (c) scs_destroy(c);
    return rc == SCS_E_IO ? -1 : 1;
}

int main(int argc, char* argv[]) {
    time_t start_t = time(NULL);
    if (argc == 1) { usage(argv[0]); return 0; }
    string subcmd = argv[1];
    if (subcmd == "-h" || subcmd == "--help") { usage(argv[0]); return 0; }
    if (subcmd == "-v" || subcmd == "--version") { cerr << "SCSsim version 1.0" << endl; return 0; }
    if (subcmd == "simuvars") return main_simuvars(argc - 1, argv + 1, start_t);
    if (subcmd == "learn") {
        cerr << "Error: subcommand \"" << subcmd << "\" is not part of the B200 build; use the reference scssim for it." << endl;
        return 1;
    }
    if (subcmd != "genreads") { cerr << "Error: unrecognized subcommand \"" << subcmd << "\"." << endl; usage(argv[0]); return 0; }

    string modelFile, inputFile, outputPrefix, layout = "PE";
    scs_params P; scs_default_params(&P);
    P.seed = (uint64_t)start_t;
    int threads = 1, gpus = 1; bool threads_given = false;
    struct option long_options[] = {{"help", no_argument, 0, 'h'},          {"input", required_argument, 0, 'i'},
                                    {"primers", required_argument, 0, 'p'}, {"gamma", required_argument, 0, 'r'},
                                    {"model", required_argument, 0, 'm'},   {"layout", required_argument, 0, 'l'},
                                    {"coverage", required_argument, 0, 'c'}, {"isize", required_argument, 0, 's'},
                                    {"threads", required_argument, 0, 't'}, {"output", required_argument, 0, 'o'},
                                    {"seed", required_argument, 0, 1000},   {"device", required_argument, 0, 1001},
                                    {"gpus", required_argument, 0, 1002},   {"gz", no_argument, 0, 1003},
                                    {0, 0, 0, 0}};
    int c;
    argc -= 1; argv += 1;
    while ((c = getopt_long(argc, argv, "hi:p:r:m:l:c:s:t:o:", long_options, NULL)) != -1) {
        switch (c) {
            case 'h': usage_genReads(argv[0]); return 0;
            case 'i': inputFile = optarg; break;
            case 'p': P.primers = atol(optarg); break;
            case 'r': P.gamma = atof(optarg); break;
            case 'm': modelFile = optarg; break;
            case 'l': layout = optarg; break;
            case 'c': P.coverage = atof(optarg); break;
            case 's': P.isize = atoi(optarg); break;
            case 't': threads = atoi(optarg); threads_given = true; break;
            case 'o': outputPrefix = optarg; break;
            case 1000: P.seed = strtoull(optarg, NULL, 0); break;
            case 1001: P.device = atoi(optarg); break;
            case 1002: gpus = atoi(optarg); break;
            case 1003: P.gzip = 1; break;
            default: usage_genReads(argv[0]); return 1;
        }
    }
    if (inputFile.empty()) { cerr << "Error: reference file (.fasta) not specified!" << endl; usage_genReads(argv[0]); return 1; }
    if (P.primers < 1000) { cerr << "Error: the value of parameter \"primers\" should be at least 1000!" << endl; return 1; }
    if (P.gamma <= 0 || P.gamma > 1e-8) { cerr << "Error: the value of parameter \"gamma\" should be in 0~1e-8!" << endl; return 1; }
    if (modelFile.empty()) { cerr << "Error: sequencing profile must be specified!" << endl; usage_genReads(argv[0]); return 1; }
    if (outputPrefix.empty()) { cerr << "Error: the prefix of output file not specified!" << endl; usage_genReads(argv[0]); return 1; }
    if (layout.empty()) { cerr << "Warning: sequence layout not specified!" << endl << "use the default value: \"PE for paired-end\"" << endl; layout = "PE"; }
    else if (layout != "SE" && layout != "PE") { cerr << "Error: sequence layout incorrectly specified!" << endl << "should be SE (single end) or PE (paired-end)" << endl; return 1; }
    if (P.coverage <= 0) { cerr << "Error: sequencing coverage not properly specified!" << endl; return 1; }
    if (threads < 1) { cerr << "Error: number of threads should be a positive integer!" << endl; return 1; }
    P.paired = layout == "PE";
    P.io_threads = threads_given ? threads : 0;   // the reference's worker threads become the host threads that write the FASTQ slabs
    if (gpus < 1) { cerr << "Error: number of GPUs should be a positive integer!" << endl; return 1; }

    if (gpus > 1) {
        string fa = inputFile;
        if (is_gz(fa)) {
            string plain;
            if (!gunzip_beside(fa, &plain)) { cerr << "could not open " << fa << endl; return 1; }
            fa = plain;
        }
        if (!getenv("SCS_CLI_SAME_DEVICE") && P.device + gpus > scs_device_count()) {
            cerr << "Error: --gpus " << gpus << " from device " << P.device << " needs " << P.device + gpus << " CUDA devices, found " << scs_device_count() << endl;
            return 1;
        }
        // Collectives: NCCL inside the library, one communicator per worker (NVLink / NVSwitch). The contexts are created first, so a
        // GPU that cannot be opened is reported before anybody waits in ncclCommInitRank. SCS_CLI_SAME_DEVICE=1 (test hook: all
        // workers on one GPU, which NCCL refuses) uses host sums between the worker threads instead.
        const bool same_device = getenv("SCS_CLI_SAME_DEVICE") != nullptr;
        ThreadSum coll(gpus);
        std::vector<int> rcs(gpus, 0); std::vector<string> errs(gpus);
        std::vector<uint64_t> reads(gpus, 0);
        std::vector<scs_ctx*> ctxs(gpus, nullptr);
        char nccl_id[SCS_NCCL_ID_BYTES];
        if (!same_device && scs_nccl_unique_id(nccl_id)) { cerr << scs_last_error(nullptr) << endl; return 1; }
        for (int r = 0; r < gpus; r++) {
            scs_params Q = P; Q.rank = r; Q.world = gpus; Q.balance = 1;   // contiguous slot ranges; the shards are consecutive regions of the output files
            Q.device = same_device ? P.device : P.device + r;
            if (int rc = scs_create(&Q, &ctxs[r])) {
                cerr << scs_last_error(nullptr) << endl;
                for (scs_ctx* c : ctxs) if (c) scs_destroy(c);
                return rc == SCS_E_IO ? -1 : 1;
            }
        }
        cerr << "\nReference sequence and profile are loaded by " << gpus << " GPU workers" << endl << "\nMALBAC amplification..." << endl;
        std::mutex abort_mu;
        auto abort_all = [&](int self) {   // a failed worker releases everybody who waits for it in a collective
            std::lock_guard<std::mutex> lk(abort_mu);
            coll.abort();
            if (!same_device) for (int r = 0; r < gpus; r++) if (r != self) scs_nccl_abort(ctxs[r]);
        };
        auto worker = [&](int r) {
            scs_ctx* c = ctxs[r];
            int rc = 0;
            if (same_device) scs_set_collectives(c, sum_u64, sum_f64, &coll);
            else rc = scs_nccl_init(c, nccl_id);
            if (rc || (rc = scs_load_genome(c, fa.c_str())) || (rc = scs_load_profile(c, modelFile.c_str())) || (rc = scs_create_frags(c)) ||
                (rc = scs_amplify(c)) || (rc = scs_set_read_counts(c)) || (rc = scs_yield_reads(c, outputPrefix.c_str()))) {
                errs[r] = scs_last_error(c); rcs[r] = rc; abort_all(r);
            }
            scs_stats st; scs_get_stats(c, &st); reads[r] = st.reads_requested;
        };
        std::vector<std::thread> ts;
        for (int r = 0; r < gpus; r++) ts.emplace_back(worker, r);
        for (auto& t : ts) t.join();
        for (scs_ctx* c : ctxs) scs_destroy(c);
        // report the worker that failed first-hand, not the ones that were released from a collective by its failure
        int bad = -1;
        for (int r = 0; r < gpus; r++)
            if (rcs[r] && (bad < 0 || errs[bad].find("callback failed") != string::npos || errs[bad].find("NCCL error") != string::npos)) bad = r;
        if (bad >= 0) { cerr << errs[bad] << endl; return rcs[bad] == SCS_E_IO ? -1 : 1; }
        // every worker wrote its shard at its final offset of the reference's file names (scs_yield_reads: sizing pass + exchange)
        cerr << "\nNumber of reads to generate: " << reads[0] << endl << "\n*****Producing reads*****" << endl;
        cerr << "\nReads generation done!" << endl;
        long used = (long)(time(NULL) - start_t);
        cerr << "\nElapsed time: " << used / 60 << " minutes and " << used % 60 << " seconds!\n" << endl;
        return 0;
    }

    scs_ctx* ctx = nullptr;
    int rc = scs_create(&P, &ctx);
    if (rc) return die(nullptr, rc);
    // .gz input: the reference gunzips beside the input (lib/genome/Genome.cpp:183-187)
    string fa = inputFile;
    if (is_gz(fa)) {
        string plain;
        if (!gunzip_beside(fa, &plain)) { cerr << "could not open " << fa << endl; scs_destroy(ctx); return 1; }
        fa = plain;
    }
    if ((rc = scs_load_genome(ctx, fa.c_str()))) return die(ctx, rc);
    cerr << "\nReference sequence was loaded from file " << inputFile << endl;
    if ((rc = scs_load_profile(ctx, modelFile.c_str()))) return die(ctx, rc);
    cerr << "profile was loaded from file " << modelFile << endl;
    if ((rc = scs_create_frags(ctx))) return die(ctx, rc);
    cerr << "\nMALBAC amplification..." << endl;
    if ((rc = scs_amplify(ctx))) return die(ctx, rc);
    if ((rc = scs_set_read_counts(ctx))) return die(ctx, rc);
    scs_stats st; scs_get_stats(ctx, &st);
    cerr << "\nNumber of reads to generate: " << st.reads_requested << endl;
    cerr << "\n*****Producing reads*****" << endl;
    if ((rc = scs_yield_reads(ctx, outputPrefix.c_str()))) return die(ctx, rc);
    cerr << "\nReads generation done!" << endl;
    scs_destroy(ctx);
    long used = (long)(time(NULL) - start_t);
    cerr << "\nElapsed time: " << used / 60 << " minutes and " << used % 60 << " seconds!\n" << endl;
    return 0;
}
