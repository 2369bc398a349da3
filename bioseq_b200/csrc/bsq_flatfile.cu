// FlatFile: the packed on-disk sequence store of the reference (src/fxstats.cpp:26-134)
// as a feeder of the GPU path.  The file layout
//     uint64 nseqs | uint64 offsets[nseqs+1] | residue bytes        (src/fxstats.cpp:50-59)
// already is the bytes + offsets form the kernels read, so a range of sequences goes from
// the mapping (or from a pinned in-memory copy of the file) to the device without any
// per-sequence host work; see bsq_tokenize_host in bsq_host.cu.
//
// The FASTA/FASTQ reader below is written against the record rules of the reference's
// vendored kseq.h (src/kseq.h:173-216, which is what FlatFile::make drives at
// src/fxstats.cpp:40-49) -- the rules are restated at parse_fastx() -- not against its code.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bsq_internal.h"

using bsq::fail;

namespace {

// Whole input, decompressed (gzread passes plain files through unchanged).
int slurp(const char *path, std::vector<uint8_t> &data) {
    gzFile fp = gzopen(path, "r");
    if (fp == nullptr) return fail(BSQ_ERR_IO, std::string(path) + " failed to open");  // src/fxstats.cpp:41
    gzbuffer(fp, 1 << 20);
    size_t used = 0;
    data.resize(size_t(4) << 20);
    for (;;) {
        if (data.size() - used < (size_t(1) << 20)) data.resize(data.size() * 2);
        const int got = gzread(fp, data.data() + used, static_cast<unsigned>(std::min<size_t>(data.size() - used, size_t(1) << 30)));
        if (got < 0) {
            int errnum = 0;
            const std::string why = gzerror(fp, &errnum);
            gzclose(fp);
            return fail(BSQ_ERR_IO, std::string(path) + ": read error: " + why);
        }
        if (got == 0) break;
        used += static_cast<size_t>(got);
    }
    gzclose(fp);
    data.resize(used);
    return BSQ_OK;
}

// Record rules (kseq.h as vendored by the reference, src/kseq.h:173-216):
//  * a record starts at the next '>' or '@' -- searched byte-wise, not only at line starts, when
//    the previous record ended inside a FASTQ quality block; otherwise the header character is
//    the one that terminated the previous record's sequence lines;
//  * name = up to the first whitespace; if that whitespace is not '\n' the rest of the line is a
//    comment; nothing left after the header character ends the file;
//  * sequence = following lines concatenated until a line that starts with '>', '+' or '@' (or
//    EOF); empty lines are skipped; after each line one trailing '\r' is dropped when the
//    accumulated sequence is longer than one character;
//  * '+' starts a FASTQ quality block: skip the rest of that line, then append lines until the
//    quality is at least as long as the sequence; a missing or differently sized quality string
//    ends the parse (the record is not counted, like `while(kseq_read(ks) >= 0)` at
//    src/fxstats.cpp:45 stopping on -2).
// `residues` receives the concatenated sequences, `offsets` their boundaries (leading 0).
int parse_fastx(const std::vector<uint8_t> &data, std::vector<uint8_t> *residues, std::vector<uint64_t> &offsets) {
    const size_t n = data.size();
    const uint8_t *d = data.data();
    size_t p = 0;
    auto getc = [&]() -> int { return p < n ? d[p++] : -1; };
    auto line_end = [&](size_t from) -> size_t {
        const void *nl = from < n ? std::memchr(d + from, '\n', n - from) : nullptr;
        return nl ? static_cast<size_t>(static_cast<const uint8_t *>(nl) - d) : n;
    };
    offsets.assign(1, 0);
    std::vector<uint8_t> seq;
    std::string qual;
    int last_char = 0, c;
    for (;;) {
        if (last_char == 0) {
            while ((c = getc()) >= 0 && c != '>' && c != '@') {}
            if (c < 0) break;
            last_char = c;
        }
        if (p >= n) break;  // header character at the very end: no record
        size_t i = p;
        while (i < n && !std::isspace(d[i])) ++i;
        const int delim = i < n ? d[i] : 0;
        p = i < n ? i + 1 : n;
        if (delim != '\n') {
            const size_t e = line_end(p);
            p = e < n ? e + 1 : n;
        }
        seq.clear();
        while ((c = getc()) >= 0 && c != '>' && c != '+' && c != '@') {
            if (c == '\n') continue;
            seq.push_back(static_cast<uint8_t>(c));
            if (p < n) {  // rest of the line
                const size_t e = line_end(p);
                seq.insert(seq.end(), d + p, d + e);
                p = e < n ? e + 1 : n;
                if (seq.size() > 1 && seq.back() == '\r') seq.pop_back();
            }
        }
        if (c == '>' || c == '@') last_char = c;
        if (c == '+') {
            while ((c = getc()) >= 0 && c != '\n') {}
            if (c < 0) break;  // no quality string
            qual.clear();
            while (p < n) {
                const size_t e = line_end(p);
                qual.append(reinterpret_cast<const char *>(d + p), e - p);
                p = e < n ? e + 1 : n;
                if (qual.size() > 1 && qual.back() == '\r') qual.pop_back();
                if (qual.size() >= seq.size()) break;
            }
            last_char = 0;
            if (qual.size() != seq.size()) break;  // truncated / oversized quality
        }
        if (seq.size() > 0xFFFFFFFFull)
            return fail(BSQ_ERR_ARG, "Cannot handle sequences longer than 2^32 - 1");  // src/fxstats.cpp:46
        if (residues) residues->insert(residues->end(), seq.begin(), seq.end());
        offsets.push_back(offsets.back() + seq.size());
    }
    return BSQ_OK;
}

}  // namespace

struct bsq_flatfile {
    std::string path;
    uint8_t *base = nullptr;  // whole file: header + offsets + residues
    size_t size = 0;
    int mode = BSQ_FF_MMAP;
    int64_t nseqs = 0, seq_offset = 0, max_seq_len = 0;
};

extern "C" {

int bsq_flatfile_make(const char *inpath, const char *outpath, int64_t *nseqs, int64_t *max_seq_len) {
    if (inpath == nullptr) return fail(BSQ_ERR_ARG, "null path");
    const std::string out = (outpath == nullptr || outpath[0] == '\0') ? std::string(inpath) + ".ff" : std::string(outpath);
    std::vector<uint8_t> data, residues;
    std::vector<uint64_t> offsets;
    if (int rc = slurp(inpath, data)) return rc;
    residues.reserve(data.size());
    if (int rc = parse_fastx(data, &residues, offsets)) return rc;
    std::vector<uint8_t>().swap(data);
    std::FILE *ofp = std::fopen(out.c_str(), "w");
    if (ofp == nullptr) return fail(BSQ_ERR_IO, out + " could not be opened for writing");  // src/fxstats.cpp:53
    const uint64_t n = offsets.size() - 1;
    bool ok = std::fwrite(&n, sizeof(n), 1, ofp) == 1;
    ok = ok && std::fwrite(offsets.data(), sizeof(uint64_t), offsets.size(), ofp) == offsets.size();
    ok = ok && (residues.empty() || std::fwrite(residues.data(), 1, residues.size(), ofp) == residues.size());
    ok = (std::fclose(ofp) == 0) && ok;
    if (!ok) return fail(BSQ_ERR_IO, out + ": write failed");
    uint64_t longest = 0;
    for (size_t i = 0; i + 1 < offsets.size(); ++i) longest = std::max(longest, offsets[i + 1] - offsets[i]);
    if (nseqs) *nseqs = static_cast<int64_t>(n);
    if (max_seq_len) *max_seq_len = static_cast<int64_t>(longest);
    return BSQ_OK;
}

int bsq_flatfile_open(bsq_flatfile **outp, const char *path, int64_t maxseqlen, int mode) {
    if (outp == nullptr || path == nullptr) return fail(BSQ_ERR_ARG, "null argument");
    if (mode != BSQ_FF_MMAP && mode != BSQ_FF_PINNED && mode != BSQ_FF_MMAP_PREFAULT && mode != BSQ_FF_MMAP_REGISTERED)
        return fail(BSQ_ERR_ARG, "bad FlatFile mode");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(BSQ_ERR_IO, std::strerror(errno));  // the reference surfaces mio's system_error text
    struct stat sb;
    if (::fstat(fd, &sb) != 0) {
        const int e = errno;
        ::close(fd);
        return fail(BSQ_ERR_IO, std::strerror(e));
    }
    const size_t size = static_cast<size_t>(sb.st_size);
    if (size < 16) {
        ::close(fd);
        return fail(BSQ_ERR_IO, std::string(path) + ": not a FlatFile (shorter than its header)");
    }
    uint8_t *base = nullptr;
    bool registered = false;
    if (mode != BSQ_FF_PINNED) {
        int flags = MAP_SHARED;
#ifdef MAP_POPULATE
        if (mode != BSQ_FF_MMAP) flags |= MAP_POPULATE;  // page tables filled now, not fault by fault under the first pass
#endif
        void *m = ::mmap(nullptr, size, PROT_READ, flags, fd, 0);
        const int e = errno;
        ::close(fd);
        if (m == MAP_FAILED) return fail(BSQ_ERR_IO, std::strerror(e));
        base = static_cast<uint8_t *>(m);
        if (mode == BSQ_FF_MMAP_REGISTERED) {
            // page-lock the mapping where it lies: the DMA engine then reads the page cache directly
            registered = cudaHostRegister(base, size, cudaHostRegisterPortable | cudaHostRegisterReadOnly) == cudaSuccess;
            if (!registered) {
                // Platforms without read-only registration pin pages for writing, which a PROT_READ mapping refuses:
                // map the file shared and writable instead (nothing is ever written through it) when it may be.
                cudaGetLastError();
                const int wfd = ::open(path, O_RDWR);
                if (wfd >= 0) {
                    void *mw = ::mmap(nullptr, size, PROT_READ | PROT_WRITE, flags, wfd, 0);
                    ::close(wfd);
                    if (mw != MAP_FAILED) {
                        if (cudaHostRegister(mw, size, cudaHostRegisterPortable) == cudaSuccess) {
                            ::munmap(base, size);
                            base = static_cast<uint8_t *>(mw);
                            registered = true;
                        } else {
                            cudaGetLastError();
                            ::munmap(mw, size);
                        }
                    }
                }
            }
            if (!registered) mode = BSQ_FF_MMAP_PREFAULT;  // (not an error: the file stays a prefaulted pageable mapping)
        }
    } else {
        // +32: the kernels' host-side staging copies whole ranges; keep a readable tail like the pack layer
        cudaError_t ce = cudaHostAlloc(reinterpret_cast<void **>(&base), size + 32, cudaHostAllocPortable);
        if (ce != cudaSuccess) {
            ::close(fd);
            return fail(BSQ_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(ce));
        }
        size_t got = 0;
        while (got < size) {
            const ssize_t r = ::pread(fd, base + got, std::min<size_t>(size - got, size_t(1) << 30), static_cast<off_t>(got));
            if (r <= 0) {
                const int e = errno;
                ::close(fd);
                cudaFreeHost(base);
                return fail(BSQ_ERR_IO, r == 0 ? std::string(path) + ": unexpected end of file" : std::string(std::strerror(e)));
            }
            got += static_cast<size_t>(r);
        }
        std::memset(base + size, 0, 32);
        ::close(fd);
    }
    auto release = [&]() {
        if (registered) cudaHostUnregister(base);
        if (mode != BSQ_FF_PINNED) ::munmap(base, size);
        else cudaFreeHost(base);
    };
    uint64_t n;
    std::memcpy(&n, base, 8);
    const uint64_t *offs = reinterpret_cast<const uint64_t *>(base) + 1;
    if (n > (size - 16) / 8) {
        release();
        return fail(BSQ_ERR_IO, std::string(path) + ": not a FlatFile (offset table runs past the end of the file)");
    }
    const size_t seq_offset = (static_cast<size_t>(n) + 2) * 8;  // src/fxstats.cpp:67
    bool sane = offs[0] == 0 && offs[n] <= size - seq_offset;
    uint64_t longest = 0;
    for (uint64_t i = 0; sane && i < n; ++i) {
        sane = offs[i] <= offs[i + 1];
        longest = std::max(longest, offs[i + 1] - offs[i]);
    }
    if (!sane) {
        release();
        return fail(BSQ_ERR_IO, std::string(path) + ": not a FlatFile (offsets are not a prefix sum inside the file)");
    }
    bsq_flatfile *f = new bsq_flatfile();
    f->path = path;
    f->base = base;
    f->size = size;
    f->mode = mode;
    f->nseqs = static_cast<int64_t>(n);
    f->seq_offset = static_cast<int64_t>(seq_offset);
    // src/fxstats.cpp:69-74: the caller's value is trusted when given; lengths are kept in 32 bits
    f->max_seq_len = maxseqlen >= 0 ? static_cast<int64_t>(static_cast<uint32_t>(maxseqlen))
                                    : static_cast<int64_t>(static_cast<uint32_t>(longest));
    *outp = f;
    return BSQ_OK;
}

void bsq_flatfile_close(bsq_flatfile *f) {
    if (f == nullptr) return;
    if (f->base != nullptr) {
        if (f->mode == BSQ_FF_MMAP_REGISTERED) cudaHostUnregister(f->base);
        if (f->mode != BSQ_FF_PINNED) ::munmap(f->base, f->size);
        else cudaFreeHost(f->base);
    }
    delete f;
}

int64_t bsq_flatfile_nseqs(const bsq_flatfile *f) { return f ? f->nseqs : 0; }
int64_t bsq_flatfile_seq_offset(const bsq_flatfile *f) { return f ? f->seq_offset : 0; }
int64_t bsq_flatfile_max_seq_len(const bsq_flatfile *f) { return f ? f->max_seq_len : 0; }
const int64_t *bsq_flatfile_offsets(const bsq_flatfile *f) { return f ? reinterpret_cast<const int64_t *>(f->base + 8) : nullptr; }
const uint8_t *bsq_flatfile_bytes(const bsq_flatfile *f) { return f ? f->base + f->seq_offset : nullptr; }
int bsq_flatfile_is_pinned(const bsq_flatfile *f) { return f && (f->mode == BSQ_FF_PINNED || f->mode == BSQ_FF_MMAP_REGISTERED); }

int bsq_fastx_lengths(const char *path, int64_t **lens, int64_t *n) {
    if (path == nullptr || lens == nullptr || n == nullptr) return fail(BSQ_ERR_ARG, "null argument");
    std::vector<uint8_t> data;
    std::vector<uint64_t> offsets;
    if (int rc = slurp(path, data)) return rc;
    if (int rc = parse_fastx(data, nullptr, offsets)) return rc;
    const size_t cnt = offsets.size() - 1;
    int64_t *out = static_cast<int64_t *>(std::malloc(std::max<size_t>(cnt, 1) * sizeof(int64_t)));
    if (out == nullptr) return fail(BSQ_ERR_NOMEM, "out of host memory");
    for (size_t i = 0; i < cnt; ++i) out[i] = static_cast<int64_t>(offsets[i + 1] - offsets[i]);
    *lens = out;
    *n = static_cast<int64_t>(cnt);
    return BSQ_OK;
}

void bsq_free(void *p) { std::free(p); }

}  // extern "C"
