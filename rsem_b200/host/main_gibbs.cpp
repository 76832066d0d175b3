// rsem-run-gibbs, B200 edition: same argv and files as the reference (/root/reference/Gibbs.cpp:425-530).
//
//   rsem-run-gibbs reference_name imdName statName BURNIN NSAMPLES GAP [-p #Threads] [--seed seed]
//                  [--pseudo-count pseudo_count] [--prior file] [-q]
//
// -p keeps its meaning: it is the number of independent chains (the results depend on it, exactly as in
// the reference: NSAMPLES is split over the chains and each chain gets its own mt19937 seed).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <ctime>
#include <future>
#include <set>
#include <thread>

#include "host.hpp"

using namespace host;

namespace {

}  // namespace

int main(int argc, char* argv[]) {
    if (argc < 7) {
        printf("Usage: rsem-run-gibbs reference_name imdName statName BURNIN NSAMPLES GAP [-p #Threads] [--seed seed] [--pseudo-count pseudo_count] [--prior file] [-q]\n");
        printf("\n");
        printf("Format of the prior file:\n");
        printf("- One isoform's prior per line\n");
        printf("- Priors must be in the same order as in the .ti file\n");
        printf("- Priors for those to-be-omitted isoforms must be included as well\n");
        printf("- Comments can be added after prior separated by space(s)\n");
        exit(-1);
    }
    const std::string refName = argv[1], imdName = argv[2], statName = argv[3];
    const int BURNIN = atoi(argv[4]), NSAMPLES = atoi(argv[5]), GAP = atoi(argv[6]);
    int nThreads = 1;
    bool hasSeed = false, quiet = false, has_prior = false;
    uint32_t seed = 0;
    double pseudoC = 1.0;
    std::string fprior;
    for (int i = 7; i < argc; ++i) {
        if (!strcmp(argv[i], "-p") && i + 1 < argc) nThreads = atoi(argv[i + 1]);
        if (!strcmp(argv[i], "--seed") && i + 1 < argc) { hasSeed = true; seed = parse_seed(argv[i + 1]); }
        if (!strcmp(argv[i], "--pseudo-count") && i + 1 < argc) pseudoC = atof(argv[i + 1]);
        if (!strcmp(argv[i], "-q")) quiet = true;
        if (!strcmp(argv[i], "--prior") && i + 1 < argc) { has_prior = true; fprior = argv[i + 1]; }
    }
    g_verbose = !quiet;
    if (!(NSAMPLES > 1)) die("NSAMPLES must be larger than 1, otherwise the posterior variance cannot be calculated!");
    if (GAP < 1) die("GAP must be at least 1!");
    if (nThreads > NSAMPLES) {
        nThreads = NSAMPLES;
        fprintf(stderr, "Warning: Number of samples is less than number of threads! Change the number of threads to %d!\n", nThreads);
    }
    if (nThreads < 1) nThreads = 1;
    {   // -p means chains here (as in the reference); file parsing uses the host's cores
        const char* e = getenv("RSEM_B200_IO_THREADS");
        const unsigned hw = std::thread::hardware_concurrency();
        g_io_threads = e ? atoi(e) : (int)std::min<unsigned>(hw ? hw : 1, 32);
    }

    // CUDA context creation takes seconds on a multi-GPU node: overlap it with the .ofg parsing
    struct CtxResult { int rc; rsem_b200_ctx* ctx; std::string err; };
    std::future<CtxResult> ctx_future = std::async(std::launch::async, []() {
        CtxResult r{0, nullptr, ""};
        const char* dev = getenv("RSEM_B200_DEVICE");
        r.rc = rsem_b200_ctx_create(dev ? atoi(dev) : 0, &r.ctx);
        if (r.rc != 0) r.err = rsem_b200_last_error();
        return r;
    });
    const bool timing = getenv("RSEM_B200_TIMING") != nullptr;
    auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();

    // load_data
    RefData refs;
    load_refs(refName + ".seq", false, refs);
    const int M = refs.M;
    uint64_t N0 = 0;
    std::vector<uint64_t> row_ptr;
    std::vector<int32_t> sid;
    std::vector<double> conprb;
    load_ofg(imdName + ".ofg", M, N0, row_ptr, sid, conprb);
    const uint64_t N1 = row_ptr.size() - 1;
    if (g_verbose) printf("Loading data is finished!\n");

    std::vector<int> gi, gt, ta;
    load_groups(refName + ".grp", gi);
    const int m = (int)gi.size() - 1;
    const bool alleleS = load_allele_groups(refName, gt, ta);
    if (g_verbose) printf("Loading group information is finished!\n");

    // load_omit_info (Gibbs.cpp:152-167)
    std::vector<int32_t> init_counts(M + 1, 0);
    double totc = M + 1;
    {
        FILE* fi = fopen((imdName + ".omit").c_str(), "r");
        if (!fi) die("Cannot open " + imdName + ".omit!");
        int tid;
        while (fscanf(fi, "%d", &tid) == 1) {
            if (tid < 0 || tid > M) die("Invalid transcript id in " + imdName + ".omit!");
            init_counts[tid] = -1;
            --totc;
        }
        fclose(fi);
    }
    totc = totc * pseudoC + N0 + N1;
    std::vector<double> alpha(M + 1, pseudoC);
    if (has_prior) {  // load_prior_info (Gibbs.cpp:171-194)
        alpha.assign(M + 1, 0.0);
        // one prior per transcript and line, anything after the number is a comment (Gibbs.cpp:178-184 uses getline:
        // no limit on the line length)
        std::vector<char> buf = slurp(fprior);
        buf.push_back('\0');
        const char* q = buf.data();
        const char* const end = q + buf.size() - 1;
        for (int i = 1; i <= M; ++i) {
            if (q >= end) die("The prior file " + fprior + " has fewer lines than transcripts!");
            const char* eol = static_cast<const char*>(memchr(q, '\n', (size_t)(end - q)));
            if (!eol) eol = end;
            char* after = nullptr;
            const double prior = strtod(q, &after);
            if (after == q || after > eol) die("Cannot read the prior of transcript " + std::to_string(i) + " from " + fprior + "!");
            if (init_counts[i] == 0) alpha[i] = prior;
            q = eol < end ? eol + 1 : end;
        }
        totc = 1;
        for (int i = 1; i <= M; ++i)
            if (init_counts[i] == 0) totc += alpha[i];
        totc += N0 + N1;
    }

    // init_model_related: eel from the .model's gld, mw copy
    int model_type = 0;
    LenDistH gld;
    std::vector<double> mw, eel;
    HostModel::read_for_gibbs(statName + ".model", M, model_type, gld, mw);
    calc_eel(refs, gld, eel);

    if (g_verbose) printf("Gibbs started!\n");

    // init(): samples per chain and per-chain engines (Gibbs.cpp:207-254)
    std::vector<int32_t> chain_samples(nThreads);
    std::vector<uint32_t> chain_seeds(nThreads);
    {
        const int quotient = NSAMPLES / nThreads, left = NSAMPLES % nThreads;
        Mt seedEngine(hasSeed ? seed : (uint32_t)time(NULL));
        std::set<uint32_t> used;
        for (int i = 0; i < nThreads; ++i) {
            chain_samples[i] = quotient + (i < left ? 1 : 0);
            uint32_t s;
            do { s = seedEngine.next(); } while (used.count(s));
            used.insert(s);
            chain_seeds[i] = s;
        }
    }
    if (g_verbose) printf("Initialization finished!\n");

    rsem_b200_ctx* ctx = nullptr;
    {
        CtxResult r = ctx_future.get();
        if (r.rc != 0) die("rsem_b200: ctx_create failed: " + r.err);
        ctx = r.ctx;
    }
    const double t_loaded = now();
    check_rc(rsem_b200_gibbs_upload(ctx, N1, sid.size(), M, row_ptr.data(), sid.data(), conprb.data()), "gibbs_upload");
    const double t_uploaded = now();

    rsem_b200_gibbs_params gp;
    memset(&gp, 0, sizeof gp);
    gp.M = M; gp.burnin = BURNIN; gp.gap = GAP; gp.n_chains = nThreads;
    gp.chain_samples = chain_samples.data();
    gp.chain_seeds = chain_seeds.data();
    gp.n0 = (double)N0;
    gp.init_counts = init_counts.data();
    gp.pseudo_counts = alpha.data();
    gp.totc = totc;
    gp.eel = eel.data();
    gp.mw = mw.data();
    gp.n_genes = m;
    gp.gene_start = gi.data();
    std::vector<int32_t> cvs((size_t)NSAMPLES * (M + 1));
    std::vector<double> sum_c(M + 1), sum_c2(M + 1), sum_tpm(M + 1), sum_fpkm(M + 1), sum_g2(m);
    rsem_b200_gibbs_out go;
    go.count_vectors = cvs.data();
    go.sum_c = sum_c.data(); go.sum_c2 = sum_c2.data(); go.sum_tpm = sum_tpm.data(); go.sum_fpkm = sum_fpkm.data();
    go.sum_gene_c2 = sum_g2.data();
    check_rc(rsem_b200_gibbs_run(ctx, &gp, &go), "gibbs_run");
    if (timing)
        fprintf(stderr, "rsem-run-gibbs timing: load+ctx %.3f s, upload+components %.3f s, sampling %.3f s (%d chains)\n",
                t_loaded - t_begin, t_uploaded - t_loaded, now() - t_uploaded, nThreads);
    rsem_b200_ctx_destroy(ctx);

    // imd.countvectors<t> (Gibbs.cpp:257-262)
    {
        size_t at = 0;
        static char buf[1 << 20];
        for (int t = 0; t < nThreads; ++t) {
            FILE* fo = fopen((imdName + ".countvectors" + std::to_string(t)).c_str(), "w");
            if (!fo) die("Cannot open " + imdName + ".countvectors" + std::to_string(t) + " for writing!");
            setvbuf(fo, buf, _IOFBF, sizeof buf);
            for (int s = 0; s < chain_samples[t]; ++s, ++at) {
                const int32_t* c = &cvs[at * (M + 1)];
                for (int i = 0; i < M; ++i) fprintf(fo, "%d ", c[i]);
                fprintf(fo, "%d\n", c[M]);
            }
            fclose(fo);
        }
    }

    // release(): means and unbiased variances (Gibbs.cpp:399-422)
    std::vector<double> pme_c(M + 1), pve_c(M + 1), pme_tpm(M + 1), pme_fpkm(M + 1), pve_g(m), pve_t;
    for (int i = 0; i <= M; ++i) {
        pme_c[i] = sum_c[i] / NSAMPLES;
        pve_c[i] = (sum_c2[i] - double(NSAMPLES) * pme_c[i] * pme_c[i]) / double(NSAMPLES - 1);
        if (pve_c[i] < 0.0) pve_c[i] = 0.0;
        pme_tpm[i] = sum_tpm[i] / NSAMPLES;
        pme_fpkm[i] = sum_fpkm[i] / NSAMPLES;
    }
    for (int i = 0; i < m; ++i) {
        double g = 0.0;
        for (int j = gi[i]; j < gi[i + 1]; ++j) g += pme_c[j];
        pve_g[i] = (sum_g2[i] - double(NSAMPLES) * g * g) / double(NSAMPLES - 1);
        if (pve_g[i] < 0.0) pve_g[i] = 0.0;
    }
    if (alleleS) {  // allele-level variances from the kept count vectors (Gibbs.cpp:338-345, 414-421)
        const int m_trans = (int)ta.size() - 1;
        pve_t.assign(m_trans, 0.0);
        for (int s = 0; s < NSAMPLES; ++s) {
            const int32_t* c = &cvs[(size_t)s * (M + 1)];
            for (int i = 0; i < m_trans; ++i) {
                double x = 0.0;
                for (int j = ta[i]; j < ta[i + 1]; ++j) x += c[j];
                pve_t[i] += x * x;
            }
        }
        for (int i = 0; i < m_trans; ++i) {
            double x = 0.0;
            for (int j = ta[i]; j < ta[i + 1]; ++j) x += pme_c[j];
            pve_t[i] = (pve_t[i] - double(NSAMPLES) * x * x) / double(NSAMPLES - 1);
            if (pve_t[i] < 0.0) pve_t[i] = 0.0;
        }
    }
    if (g_verbose) printf("Gibbs finished!\n");
    write_results_gibbs(refName, imdName, M, pme_c, pme_fpkm, pme_tpm, pve_c, pve_g, pve_t);
    return 0;
}
