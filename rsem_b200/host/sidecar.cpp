// Binary side-car `imdName.b200`: the ingest hand-off between bin/rsem-parse-alignments and bin/rsem-run-em
// (SURVEY.md 8(f).2).  What it short-cuts: the reference's `istream >>` parse of imd.dat (HitContainer.h:62-79,
// EM.cpp:127-132) and its per-round re-parse of the read files (EM.cpp:195-202).  The text files remain the contract
// (they are always written, byte-identical to the reference's); the side-car is the same content in upload layout:
//
//   header   magic "RSEMB200", version, read_type, N[3], H, byte sizes of imd.dat and of the six read files, a hash of the
//            first and last MiB of imd.dat, the short-read length threshold
//   hits     row_ptr u64[N1 + 1], sid i32[H] (sign = strand), pos i32[H], insertL i32[H] (paired-end only)
//   per read set (un, alignable, max), per mate: off u64[n + 1], base codes u8 (A0 C1 G2 T3 N4), phred values u8
//   per read set: (index, name) of every read with a mate shorter than kSidecarShortLen (the reference warns about
//            reads shorter than the seed length by name, SingleQModel.h:296-302)
//
// It is used only if every recorded file size (and the .dat fingerprint) still matches the text files next to it; otherwise rsem-run-em parses
// the text as before.  Little-endian, no alignment padding (sections are read with fread into their final vectors).
#include <sys/stat.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "host.hpp"

namespace host {

namespace {

constexpr char kMagic[8] = {'R', 'S', 'E', 'M', 'B', '2', '0', '0'};
constexpr uint32_t kVersion = 2;

struct Header {
    char magic[8];
    uint32_t version, read_type;
    uint64_t N[3], H;
    uint64_t dat_bytes;
    uint64_t read_bytes[3][2];
    uint32_t short_len, reserved;
    uint64_t dat_hash;   // FNV-1a over the first and the last MiB of imd.dat
};

uint64_t file_bytes(const std::string& path) {
    struct stat st;
    return stat(path.c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
}

// content fingerprint of a text file that is cheap at any size: FNV-1a over its first and last MiB
uint64_t edge_hash(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return 0;
    constexpr size_t kEdge = 1u << 20;
    std::vector<unsigned char> buf(kEdge);
    uint64_t hsh = 1469598103934665603ull;
    auto mix = [&](size_t n) { for (size_t i = 0; i < n; ++i) { hsh ^= buf[i]; hsh *= 1099511628211ull; } };
    mix(fread(buf.data(), 1, kEdge, f));
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz > (long)kEdge && fseek(f, sz - (long)kEdge, SEEK_SET) == 0) mix(fread(buf.data(), 1, kEdge, f));
    }
    fclose(f);
    return hsh;
}

void fill_sizes(const std::string& imd, int read_type, Header& h) {
    h.dat_bytes = file_bytes(imd + ".dat");
    h.dat_hash = edge_hash(imd + ".dat");
    for (int tag = 0; tag < 3; ++tag) {
        std::vector<std::string> files;
        read_type_files(imd, tag, read_type, files);
        for (int m = 0; m < 2; ++m) h.read_bytes[tag][m] = m < (int)files.size() ? file_bytes(files[m]) : 0;
    }
}

template <class T>
void put_vec(FILE* fo, const std::vector<T>& v, size_t n) {
    if (n && fwrite(v.data(), sizeof(T), n, fo) != n) die("Cannot write the binary side-car (disk full?)!");
}
template <class T>
bool get_vec(FILE* fi, std::vector<T>& v, size_t n) {
    v.resize(n);
    return n == 0 || fread(v.data(), sizeof(T), n, fi) == n;
}

// SingleRead(Q)::calc_lq (SingleReadQ.h:63-95) on base codes: A = 0, T = 3
bool single_lq_codes(const uint8_t* s, int len, bool has_polyA, int seed_len) {
    if (len < seed_len) return true;
    if (!has_polyA) return false;
    int numA = 0, numT = 0, numAO = 0, numTO = 0;
    const int threshold_1 = int(0.9 * len - 1.5 * sqrt(len * 1.0) + 0.5);
    const int threshold_2 = (kOlen - 1) / 2 + 1;
    for (int i = 0; i < len; ++i) {
        if (s[i] == 0) { ++numA; if (i < kOlen) ++numAO; }
        if (s[i] == 3) { ++numT; if (i >= len - kOlen) ++numTO; }
    }
    if (numA >= threshold_1) return numAO >= threshold_2;
    if (numT >= threshold_1) return numTO >= threshold_2;
    return false;
}

}  // namespace

bool sidecar_enabled() {
    const char* e = getenv("RSEM_B200_SIDECAR");
    return !(e && !strcmp(e, "0"));
}

void write_sidecar(const std::string& imd, const Sidecar& sc) {
    const std::string path = imd + ".b200";
    FILE* fo = fopen(path.c_str(), "wb");
    if (!fo) die("Cannot open " + path + " for writing!");
    Header h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.version = kVersion;
    h.read_type = (uint32_t)sc.read_type;
    for (int t = 0; t < 3; ++t) h.N[t] = sc.reads[t].n;
    h.H = sc.hits.H;
    h.short_len = kSidecarShortLen;
    fill_sizes(imd, sc.read_type, h);
    if (fwrite(&h, sizeof h, 1, fo) != 1) die("Cannot write the binary side-car (disk full?)!");
    const bool paired = sc.read_type >= 2, hasq = sc.read_type & 1;
    put_vec(fo, sc.hits.row_ptr, (size_t)sc.hits.N + 1);
    put_vec(fo, sc.hits.sid, (size_t)sc.hits.H);
    put_vec(fo, sc.hits.pos, (size_t)sc.hits.H);
    if (paired) put_vec(fo, sc.hits.insertL, (size_t)sc.hits.H);
    for (int t = 0; t < 3; ++t) {
        const ReadStore& rs = sc.reads[t];
        for (int m = 0; m < (paired ? 2 : 1); ++m) {
            put_vec(fo, rs.off[m], (size_t)rs.n + 1);
            const size_t nb = rs.n ? (size_t)rs.off[m][rs.n] : 0;
            put_vec(fo, rs.base[m], nb);
            if (hasq) put_vec(fo, rs.qual[m], nb);
        }
        const uint64_t ns = sc.shorts[t].size();
        fwrite(&ns, 8, 1, fo);
        for (const ShortRead& s : sc.shorts[t]) {
            const uint32_t len = (uint32_t)s.name.size();
            fwrite(&s.index, 8, 1, fo);
            fwrite(&len, 4, 1, fo);
            if (len) fwrite(s.name.data(), 1, len, fo);
        }
    }
    if (fclose(fo) != 0) die("Cannot write the binary side-car (disk full?)!");
}

bool load_sidecar(const std::string& imd, int read_type, Sidecar& out) {
    if (!sidecar_enabled()) return false;
    const std::string path = imd + ".b200";
    FILE* fi = fopen(path.c_str(), "rb");
    if (!fi) return false;
    Header h, now;
    bool ok = fread(&h, sizeof h, 1, fi) == 1 && !memcmp(h.magic, kMagic, 8) && h.version == kVersion &&
              (int)h.read_type == read_type && h.short_len == (uint32_t)kSidecarShortLen;
    if (ok) {  // does it still describe the text files next to it?
        memset(&now, 0, sizeof now);
        fill_sizes(imd, read_type, now);
        ok = now.dat_bytes == h.dat_bytes && now.dat_hash == h.dat_hash && !memcmp(now.read_bytes, h.read_bytes, sizeof h.read_bytes);
    }
    if (ok) {   // the counts of a damaged header must not size the allocations below: the hit sections alone need this much
        const uint64_t have = file_bytes(path);
        const uint64_t per_hit = read_type >= 2 ? 12 : 8;
        ok = h.H <= have / per_hit && h.N[1] <= have / 8 && h.N[0] <= have / 8 && h.N[2] <= have / 8;
    }
    if (!ok) { fclose(fi); return false; }
    const bool paired = read_type >= 2, hasq = read_type & 1;
    out = Sidecar();
    out.read_type = read_type;
    out.hits.N = h.N[1];
    out.hits.H = h.H;
    ok = get_vec(fi, out.hits.row_ptr, (size_t)h.N[1] + 1) && get_vec(fi, out.hits.sid, (size_t)h.H) &&
         get_vec(fi, out.hits.pos, (size_t)h.H) && (!paired || get_vec(fi, out.hits.insertL, (size_t)h.H));
    ok = ok && out.hits.row_ptr[0] == 0 && out.hits.row_ptr[h.N[1]] == h.H;
    for (int t = 0; ok && t < 3; ++t) {
        ReadStore& rs = out.reads[t];
        rs.n_mates = paired ? 2 : 1;
        rs.has_qual = hasq;
        rs.n = h.N[t];
        for (int m = 0; ok && m < rs.n_mates; ++m) {
            ok = get_vec(fi, rs.off[m], (size_t)rs.n + 1);
            if (!ok) break;
            const size_t nb = rs.n ? (size_t)rs.off[m][rs.n] : 0;
            ok = nb <= file_bytes(path) && get_vec(fi, rs.base[m], nb) && (!hasq || get_vec(fi, rs.qual[m], nb));
        }
        uint64_t ns = 0;
        ok = ok && fread(&ns, 8, 1, fi) == 1 && ns <= rs.n;
        out.shorts[t].resize(ok ? (size_t)ns : 0);
        for (size_t k = 0; ok && k < out.shorts[t].size(); ++k) {
            uint32_t len = 0;
            ok = fread(&out.shorts[t][k].index, 8, 1, fi) == 1 && fread(&len, 4, 1, fi) == 1 && len < (1u << 20);
            if (ok && len) {
                out.shorts[t][k].name.resize(len);
                ok = fread(&out.shorts[t][k].name[0], 1, len, fi) == len;
            }
        }
    }
    fclose(fi);
    if (!ok) {
        fprintf(stderr, "Warning: %s is truncated or corrupt; parsing the text files instead.\n", path.c_str());
        out = Sidecar();
    }
    return ok;
}

void finish_sidecar_reads(ReadStore& rs, const std::vector<ShortRead>& shorts, bool has_polyA, int seed_len,
                          std::vector<std::string>* short_names, uint64_t* n_short) {
    if (seed_len > kSidecarShortLen) die("internal: the side-car keeps names only for reads shorter than its threshold");
    rs.lowq.assign((size_t)rs.n, 0);
    const int s = rs.n_mates;
    parallel_ranges((size_t)rs.n, g_io_threads, [&](size_t b, size_t e, int) {
        for (size_t r = b; r < e; ++r) {
            const uint8_t* sp[2] = {nullptr, nullptr};
            int len[2] = {0, 0};
            for (int m = 0; m < s; ++m) {
                sp[m] = rs.base[m].data() + rs.off[m][r];
                len[m] = (int)(rs.off[m][r + 1] - rs.off[m][r]);
            }
            bool lq;   // same case analysis as parse_reads
            if (s == 1) lq = seed_len > 0 ? single_lq_codes(sp[0], len[0], has_polyA, seed_len) : false;
            else if (seed_len <= 0) lq = false;
            else if (len[0] < seed_len || len[1] < seed_len) lq = true;  // PairedEndReadQ.h:58-65
            else lq = single_lq_codes(sp[0], len[0], has_polyA, seed_len) && single_lq_codes(sp[1], len[1], has_polyA, seed_len);
            rs.lowq[r] = lq ? 1 : 0;
        }
    });
    uint64_t cnt = 0;
    for (const ShortRead& sr : shorts) {  // ascending read index, like the text path's thread-order concatenation
        const size_t r = (size_t)sr.index;
        if (r >= rs.n || !rs.lowq[r]) continue;
        bool is_short = false;
        for (int m = 0; m < s; ++m) is_short = is_short || (int)(rs.off[m][r + 1] - rs.off[m][r]) < seed_len;
        if (!is_short) continue;
        ++cnt;
        if (short_names && short_names->size() < 50) short_names->push_back(sr.name);
    }
    if (n_short) *n_short = cnt;
}

}  // namespace host
