// Parsers for the reference's intermediate files.  All of them work on a whole-file buffer with a
// hand-rolled tokenizer (the reference uses istream >>, 6e6 hits/s).
#include <sys/stat.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "host.hpp"

namespace host {

bool g_verbose = true;

void die(const std::string& msg) {
    fprintf(stderr, "%s\n", msg.c_str());
    exit(-1);
}

void check_rc(int rc, const char* what) {
    if (rc != 0) die(std::string("rsem_b200: ") + what + " failed: " + rsem_b200_last_error());
}

std::vector<char> slurp(const std::string& path, bool must_exist) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) {
        if (must_exist) die("Cannot open " + path + "! It may not exist.");
        return {};
    }
    struct stat st;
    std::vector<char> buf;
    if (fstat(fileno(f), &st) == 0 && st.st_size > 0) {
        buf.resize((size_t)st.st_size);
        size_t got = fread(buf.data(), 1, buf.size(), f);
        buf.resize(got);
    }
    fclose(f);
    return buf;
}

uint32_t parse_seed(const char* s) {
    uint32_t seed = 0;
    for (const char* p = s; *p; ++p) seed = seed * 10 + (uint32_t)(*p - '0');
    return seed;
}

namespace {

struct Cursor {
    const char* p;
    const char* end;
    explicit Cursor(const std::vector<char>& b) : p(b.data()), end(b.data() + b.size()) {}
    void skip_ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool eof() { skip_ws(); return p >= end; }
    bool next_i64(long long& v) {
        skip_ws();
        if (p >= end) return false;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; } else if (*p == '+') ++p;
        if (p >= end || *p < '0' || *p > '9') return false;
        long long x = 0;
        while (p < end && *p >= '0' && *p <= '9') x = x * 10 + (*p++ - '0');
        v = neg ? -x : x;
        return true;
    }
    bool next_f64(double& v) {
        skip_ws();
        if (p >= end) return false;
        char tmp[64];
        size_t n = 0;
        while (p + n < end && n < 63 && !(p[n] == ' ' || p[n] == '\n' || p[n] == '\t' || p[n] == '\r')) { tmp[n] = p[n]; ++n; }
        tmp[n] = 0;
        char* e = nullptr;
        v = strtod(tmp, &e);
        if (e == tmp) return false;
        p += (e - tmp);
        return true;
    }
    // rest of the current line (without the newline); false at EOF
    bool line(const char*& b, size_t& n) {
        if (p >= end) return false;
        b = p;
        const char* q = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!q) { n = (size_t)(end - p); p = end; }
        else { n = (size_t)(q - p); p = q + 1; }
        return true;
    }
};

int8_t g_code[256];
bool g_code_init = false;
void init_codes() {
    if (g_code_init) return;
    memset(g_code, -1, sizeof g_code);
    g_code[(int)'a'] = g_code[(int)'A'] = 0;
    g_code[(int)'c'] = g_code[(int)'C'] = 1;
    g_code[(int)'g'] = g_code[(int)'G'] = 2;
    g_code[(int)'t'] = g_code[(int)'T'] = 3;
    g_code[(int)'n'] = g_code[(int)'N'] = 4;
    g_code_init = true;
}

}  // namespace

// ---- ref.seq ------------------------------------------------------------------------------------
void load_refs(const std::string& path, bool with_seqs, RefData& r) {
    init_codes();
    std::vector<char> buf = slurp(path);
    Cursor c(buf);
    r = RefData();
    r.seq_off.push_back(0);
    r.full_len.push_back(0);
    r.tot_len.push_back(0);
    r.mask_off.push_back(0);
    long long fullLen, totLen;
    while (c.next_i64(fullLen)) {
        if (!c.next_i64(totLen)) break;
        const char* b; size_t n;
        c.line(b, n);                      // rest of the length line
        if (!c.line(b, n)) break;          // name
        if (!c.line(b, n)) break;          // sequence
        if (fullLen <= 0 || totLen < fullLen) die("Corrupted reference file " + path);
        r.seq_off.push_back(r.seq.size());
        if (with_seqs) {
            if ((long long)n < totLen) die("Reference sequence shorter than its declared length in " + path);
            size_t at = r.seq.size();
            r.seq.resize(at + (size_t)totLen);
            for (long long k = 0; k < totLen; ++k) {
                int8_t code = g_code[(unsigned char)b[k]];
                if (code < 0) {
                    fprintf(stderr, "Found unknown sequence letter %c at function get_base_id!\n", b[k]);
                    exit(-1);
                }
                r.seq[at + k] = (uint8_t)code;
            }
        }
        r.full_len.push_back((int32_t)fullLen);
        r.tot_len.push_back((int32_t)totLen);
        r.mask_off.push_back(r.mask_words.size());
        const int nw = (int)((fullLen - 1) / 32 + 1);
        for (int k = 0; k < nw; ++k) {
            long long w;
            if (!c.next_i64(w)) die("Corrupted mask words in " + path);
            r.mask_words.push_back((uint32_t)w);
        }
        c.line(b, n);  // end of the mask line
        ++r.M;
        r.has_polyA = r.has_polyA || fullLen < totLen;
    }
    r.seq_off.push_back(r.seq.size());
    if (g_verbose) printf("Refs.loadRefs finished!\n");
}

// ---- ref.ti ---------------------------------------------------------------------------------------
void load_transcripts(const std::string& path, std::vector<TranscriptInfo>& out) {
    std::ifstream fin(path.c_str());
    if (!fin.is_open()) { fprintf(stderr, "Cannot open %s! It may not exist.\n", path.c_str()); exit(-1); }
    int M, type;
    std::string line;
    fin >> M >> type;
    getline(fin, line);
    out.assign(M + 1, TranscriptInfo());
    for (int i = 1; i <= M; ++i) {
        TranscriptInfo& t = out[i];
        getline(fin, line);
        size_t tab = line.find('\t');
        t.transcript_id = line.substr(0, tab);
        t.transcript_name = tab == std::string::npos ? "" : line.substr(tab + 1);
        getline(fin, line);
        tab = line.find('\t');
        t.gene_id = line.substr(0, tab);
        t.gene_name = tab == std::string::npos ? "" : line.substr(tab + 1);
        getline(fin, t.seqname);
        std::string strand;
        int s;
        fin >> strand >> t.length >> s;
        for (int k = 0; k < s; ++k) { int a, b; fin >> a >> b; }
        getline(fin, line);
        getline(fin, line);  // "left"
    }
}

void load_groups(const std::string& path, std::vector<int>& starts) {
    FILE* fi = fopen(path.c_str(), "r");
    if (!fi) { fprintf(stderr, "Cannot open %s! It may not exist.\n", path.c_str()); exit(-1); }
    starts.clear();
    int pos;
    while (fscanf(fi, "%d", &pos) == 1) starts.push_back(pos);
    fclose(fi);
}

bool load_allele_groups(const std::string& ref_name, std::vector<int>& gt, std::vector<int>& ta) {
    FILE* a = fopen((ref_name + ".gt").c_str(), "r");
    FILE* b = fopen((ref_name + ".ta").c_str(), "r");
    const bool ok = a && b;
    if (a) fclose(a);
    if (b) fclose(b);
    if (ok) { load_groups(ref_name + ".gt", gt); load_groups(ref_name + ".ta", ta); }
    return ok;
}

void load_mparams(const std::string& path, ModelParamsH& p) {
    std::vector<char> buf = slurp(path);
    Cursor c(buf);
    long long v;
    double d;
    bool ok = c.next_i64(v); p.minL = (int)v;
    ok = ok && c.next_i64(v); p.maxL = (int)v;
    ok = ok && c.next_f64(d); p.probF = d;
    ok = ok && c.next_i64(v); p.estRSPD = v != 0;
    ok = ok && c.next_i64(v); p.B = (int)v;
    ok = ok && c.next_i64(v); p.mate_minL = (int)v;
    ok = ok && c.next_i64(v); p.mate_maxL = (int)v;
    ok = ok && c.next_f64(d); p.mean = d;
    ok = ok && c.next_f64(d); p.sd = d;
    ok = ok && c.next_i64(v); p.seedLen = (int)v;
    if (!ok) die("Cannot parse " + path + "!");
}

// ---- reads ----------------------------------------------------------------------------------------
void read_type_files(const std::string& imd, int tag, int read_type, std::vector<std::string>& files) {
    static const char* tags[3] = {"un", "alignable", "max"};
    const char* suffix = (read_type == 0 || read_type == 2) ? "fa" : "fq";
    files.clear();
    if (read_type < 2) files.push_back(imd + "_" + tags[tag] + "." + suffix);
    else {
        files.push_back(imd + "_" + tags[tag] + "_1." + suffix);
        files.push_back(imd + "_" + tags[tag] + "_2." + suffix);
    }
}

namespace {
// SingleRead(Q)::calc_lq, SingleReadQ.h:63-95
bool single_lq(const char* s, int len, bool has_polyA, int seed_len) {
    if (len < seed_len) return true;
    if (!has_polyA) return false;
    int numA = 0, numT = 0, numAO = 0, numTO = 0;
    const int threshold_1 = int(0.9 * len - 1.5 * sqrt(len * 1.0) + 0.5);
    const int threshold_2 = (kOlen - 1) / 2 + 1;
    for (int i = 0; i < len; ++i) {
        if (s[i] == 'A') { ++numA; if (i < kOlen) ++numAO; }
        if (s[i] == 'T') { ++numT; if (i >= len - kOlen) ++numTO; }
    }
    if (numA >= threshold_1) return numAO >= threshold_2;
    if (numT >= threshold_1) return numTO >= threshold_2;
    return false;
}
}  // namespace

void parse_reads(const std::string& imd, int tag, int read_type, bool has_polyA, int seed_len, ReadStore* keep,
                 ReadVisitor* visit) {
    init_codes();
    const int s = read_type >= 2 ? 2 : 1;
    const bool hasq = read_type & 1;
    std::vector<std::string> files;
    read_type_files(imd, tag, read_type, files);
    std::vector<char> buf[2];
    for (int m = 0; m < s; ++m) {
        FILE* f = fopen(files[m].c_str(), "rb");
        if (!f) { fprintf(stderr, "Cannot open %s! It may not exist.\n", files[m].c_str()); exit(-1); }
        fclose(f);
        buf[m] = slurp(files[m]);
    }
    Cursor cur[2] = {Cursor(buf[0]), Cursor(buf[s - 1])};
    if (keep) {
        keep->n_mates = s;
        keep->has_qual = hasq;
        keep->n = 0;
        for (int m = 0; m < 2; ++m) { keep->off[m].assign(1, 0); keep->base[m].clear(); keep->qual[m].clear(); }
        keep->lowq.clear();
        for (int m = 0; m < s; ++m) { keep->base[m].reserve(buf[m].size() / (hasq ? 2 : 1)); }
    }
    std::vector<uint8_t> tb[2], tq[2];
    std::string name;
    for (;;) {
        const char* seqp[2] = {nullptr, nullptr};
        const char* qualp[2] = {nullptr, nullptr};
        int len[2] = {0, 0};
        bool ok = true;
        for (int m = 0; m < s && ok; ++m) {
            const char* b; size_t n;
            if (!cur[m].line(b, n)) { ok = false; break; }
            if (n == 0 && cur[m].p >= cur[m].end) { ok = false; break; }
            if (n == 0 || b[0] != (hasq ? '@' : '>')) {
                fprintf(stderr, hasq ? "Read file does not look like a FASTQ file!\n" : "Read file does not look like a FASTA file!");
                exit(-1);
            }
            if (m == 0) name.assign(b + 1, n - 1);
            if (!cur[m].line(b, n)) { ok = false; break; }
            while (n > 0 && b[n - 1] == '\r') --n;
            seqp[m] = b; len[m] = (int)n;
            if (hasq) {
                if (!cur[m].line(b, n)) { ok = false; break; }
                if (n == 0 || b[0] != '+') { fprintf(stderr, "Read file does not look like a FASTQ file!\n"); exit(-1); }
                if (!cur[m].line(b, n)) { ok = false; break; }
                while (n > 0 && b[n - 1] == '\r') --n;
                qualp[m] = b;
                if ((int)n != len[m]) die("Read " + name + " has a different number of bases and quality values!");
            }
        }
        if (!ok) break;
        bool lq;
        if (s == 1) lq = seed_len > 0 ? single_lq(seqp[0], len[0], has_polyA, seed_len) : false;
        else if (seed_len <= 0) lq = false;
        else if (len[0] < seed_len || len[1] < seed_len) lq = true;  // PairedEndReadQ.h:58-65
        else lq = single_lq(seqp[0], len[0], has_polyA, seed_len) && single_lq(seqp[1], len[1], has_polyA, seed_len);
        for (int m = 0; m < s; ++m) {
            tb[m].resize(len[m]);
            tq[m].resize(hasq ? len[m] : 0);
            for (int k = 0; k < len[m]; ++k) {
                int8_t code = g_code[(unsigned char)seqp[m][k]];
                if (code < 0) {
                    if (!lq) { fprintf(stderr, "Found unknown sequence letter %c at function get_base_id!\n", seqp[m][k]); exit(-1); }
                    code = 4;
                }
                tb[m][k] = (uint8_t)code;
                if (hasq) {
                    const int q = (unsigned char)qualp[m][k];
                    if (q < 33 || q > 126) die("Read " + name + " has a quality character outside [33, 126]!");
                    tq[m][k] = (uint8_t)(q - 33);
                }
            }
        }
        if (visit) {
            const uint8_t* bp[2] = {tb[0].data(), tb[1].data()};
            const uint8_t* qp[2] = {hasq ? tq[0].data() : nullptr, hasq ? tq[1].data() : nullptr};
            visit->read(lq, s, bp, qp, len, name);
        }
        if (keep) {
            for (int m = 0; m < s; ++m) {
                keep->base[m].insert(keep->base[m].end(), tb[m].begin(), tb[m].end());
                if (hasq) keep->qual[m].insert(keep->qual[m].end(), tq[m].begin(), tq[m].end());
                keep->off[m].push_back(keep->base[m].size());
            }
            keep->lowq.push_back(lq ? 1 : 0);
            ++keep->n;
        }
    }
}

// ---- imd.dat --------------------------------------------------------------------------------------
void load_dat(const std::string& path, int read_type, uint64_t expect_n1, HitStore& h) {
    std::vector<char> buf = slurp(path);
    Cursor c(buf);
    long long n1, nh, rt;
    if (!c.next_i64(n1) || !c.next_i64(nh) || !c.next_i64(rt)) die("Cannot read alignments from .dat file!");
    if ((uint64_t)n1 != expect_n1) die("Number of alignable reads does not match!");
    if (rt != read_type) die("Data file (.dat) does not have the right read type!");
    const bool paired = read_type >= 2;
    h.N = (uint64_t)n1;
    h.H = (uint64_t)nh;
    h.row_ptr.assign(1, 0);
    h.row_ptr.reserve(h.N + 1);
    h.sid.clear(); h.pos.clear(); h.insertL.clear();
    h.sid.reserve(h.H); h.pos.reserve(h.H);
    if (paired) h.insertL.reserve(h.H);
    for (uint64_t i = 0; i < h.N; ++i) {
        long long k, a, b, l = 0;
        if (!c.next_i64(k) || k <= 0) die("Cannot read alignments from .dat file!");
        for (long long j = 0; j < k; ++j) {
            if (!c.next_i64(a) || !c.next_i64(b) || (paired && !c.next_i64(l))) die("Cannot read alignments from .dat file!");
            h.sid.push_back((int32_t)a);
            h.pos.push_back((int32_t)b);
            if (paired) h.insertL.push_back((int32_t)l);
        }
        h.row_ptr.push_back(h.sid.size());
    }
    h.H = h.sid.size();
}

}  // namespace host
