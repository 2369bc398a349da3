// cbioseq -- the drop-in Python extension.  Same module name, class name, method names,
// keyword arguments and defaults as the reference's pybind11 module
// (/root/reference/src/bioseq.cpp:6-11, src/tokenize.cpp:21-113, src/omp.cpp:43-49), but the
// batch methods run on the GPU through the C ABI in include/bsq.h and return torch CUDA
// tensors.  This file holds only host glue: walking the Python items, calling bsq_*, and
// turning decoded characters into Python strings.  torch is used from here through its
// Python API for device memory and the current stream only (no libtorch link).
//
// Deviations from the reference that a caller can observe:
//   * results are torch CUDA tensors, not numpy arrays;
//   * destchar 'B' gives torch.uint8 and 'b' torch.int8 (the reference binary returns int8
//     for both because it lower-cases first, src/tokenize.cpp:66,83; the bytes are the same);
//     l/L/q/Q give torch.int64 (reference: uint64, same bytes);
//   * a sequence longer than padlen raises instead of aborting the interpreter;
//   * numpy-array items are rejected with the reference's ValueError (src/tokenize.h:406-416).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bsq.h"

namespace py = pybind11;

namespace bsqpy {

// ------------------------------------------------------------------ errors
[[noreturn]] void raise_status(int rc, bool onehot = false) {
    const std::string msg = bsq_last_error();
    switch (rc) {
        case BSQ_ERR_ARG: throw py::value_error(msg);
        case BSQ_ERR_RANGE: throw py::index_error(msg);
        case BSQ_ERR_TOO_LONG:  // tokens: runtime_error (tokenize.h:458); one-hot: invalid_argument (:361)
            if (onehot) throw py::value_error(msg);
            throw std::runtime_error(msg);
        default: throw std::runtime_error(msg);
    }
}
inline void check(int rc, bool onehot = false) {
    if (rc != BSQ_OK) raise_status(rc, onehot);
}

// ------------------------------------------------------------------ host thread knob (src/omp.cpp)
int g_num_threads = 0;
py::ssize_t get_num_threads() {
    if (g_num_threads > 0) return g_num_threads;
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? hc : 1;
}
void set_num_threads(py::ssize_t n) {
    if (n > 0) g_num_threads = static_cast<int>(n);
}
// Host threads for the pack / walk of a batch call.  The reference's `nthreads` sizes its OpenMP team and defaults to 1
// (src/tokenize.cpp:65,82); here the GPU does the work the team did, and the argument only sizes the host-side gather.
// An explicit value > 1 is honoured; the default (1, or anything <= 1) means "not specified" and uses the module-wide
// setting (set_num_threads / Threading, src/omp.cpp:43-49; all hardware threads when unset).  libbsq caps its pool at
// half the hardware threads either way.  Results do not depend on it.
int host_threads(int nthreads) { return nthreads > 1 ? nthreads : static_cast<int>(get_num_threads()); }

struct Threading {
    explicit Threading(py::ssize_t n = -1) { set_num_threads(n); }
    py::ssize_t get() const { return get_num_threads(); }
    void set(py::ssize_t n) const { set_num_threads(n); }
};

// ------------------------------------------------------------------ torch plumbing
struct Torch {
    py::object mod, empty, device, cuda, uint8, int8, int16, int32, int64, float32, float64;
    static Torch &get() {
        static Torch *t = nullptr;  // leaked on purpose: must outlive interpreter teardown order
        if (t == nullptr) {
            t = new Torch();
            t->mod = py::module_::import("torch");
            t->empty = t->mod.attr("empty");
            t->device = t->mod.attr("device");
            t->cuda = t->mod.attr("cuda");
            t->uint8 = t->mod.attr("uint8");
            t->int8 = t->mod.attr("int8");
            t->int16 = t->mod.attr("int16");
            t->int32 = t->mod.attr("int32");
            t->int64 = t->mod.attr("int64");
            t->float32 = t->mod.attr("float32");
            t->float64 = t->mod.attr("float64");
        }
        return *t;
    }
    py::object dtype_of(char destchar, int kind) const {
        switch (kind) {
            case BSQ_I8: return destchar == 'B' ? uint8 : int8;
            case BSQ_I16: return int16;
            case BSQ_I32: return int32;
            case BSQ_I64: return int64;
            case BSQ_F32: return float32;
            default: return float64;
        }
    }
};

// Resolve the `device` keyword (None -> current CUDA device).  There is no CPU path.
int resolve_device(const py::object &device) {
    Torch &t = Torch::get();
    if (!t.cuda.attr("is_available")().cast<bool>())
        throw std::runtime_error("bioseq_b200: no CUDA device available -- this build has no CPU fallback");
    if (device.is_none()) return t.cuda.attr("current_device")().cast<int>();
    if (py::isinstance<py::int_>(device)) return device.cast<int>();
    py::object dev = t.device(device);
    if (dev.attr("type").cast<std::string>() != "cuda")
        throw py::value_error("bioseq_b200: device must be a CUDA device");
    py::object idx = dev.attr("index");
    return idx.is_none() ? t.cuda.attr("current_device")().cast<int>() : idx.cast<int>();
}

void *current_stream(int device) {
    Torch &t = Torch::get();
    return reinterpret_cast<void *>(t.cuda.attr("current_stream")(device).attr("cuda_stream").cast<uintptr_t>());
}

py::object new_tensor(const std::vector<int64_t> &shape, const py::object &dtype, int device) {
    Torch &t = Torch::get();
    return t.empty(py::cast(shape), py::arg("dtype") = dtype, py::arg("device") = t.device("cuda", device));
}

inline void *data_ptr(const py::object &tensor) {
    return reinterpret_cast<void *>(tensor.attr("data_ptr")().cast<uintptr_t>());
}

// ------------------------------------------------------------------ per-device staging context
struct DeviceCtx {
    std::mutex mu;
    bsq_stager *stager = nullptr;
    bsq_pack *pack = nullptr;
};
DeviceCtx &device_ctx(int device) {
    static std::mutex map_mu;
    static std::map<int, DeviceCtx *> *ctxs = new std::map<int, DeviceCtx *>();  // leaked: CUDA teardown order
    std::lock_guard<std::mutex> g(map_mu);
    auto it = ctxs->find(device);
    if (it != ctxs->end()) return *it->second;
    // built completely before it is published: a failed create (transient CUDA error, no pinned memory left)
    // must not leave a half-made context behind for every later call on this device
    std::unique_ptr<DeviceCtx> c(new DeviceCtx());
    check(bsq_stager_create(&c->stager, device));
    const int rc = bsq_pack_create(&c->pack, /*pinned=*/1);
    if (rc != BSQ_OK) {
        bsq_stager_destroy(c->stager);
        check(rc);
    }
    DeviceCtx *raw = c.release();
    (*ctxs)[device] = raw;
    return *raw;
}

// ------------------------------------------------------------------ item unpacking (src/tokenize.h:389-419)
struct Unpacked {
    py::object keepalive;  // the PySequence_Fast view: holds the items (and so their buffers) alive
    std::vector<const void *> ptrs;
    std::vector<int64_t> lens;
};

// One item -> borrowed pointer + length.  Returns false for a type the reference rejects.
inline bool unpack_one(PyObject *it, const void *&ptr, int64_t &len) {
    if (PyUnicode_Check(it)) {
        Py_ssize_t size;
        const char *s = PyUnicode_AsUTF8AndSize(it, &size);
        if (s == nullptr) throw py::error_already_set();
        ptr = s;
        len = size;
    } else if (PyBytes_Check(it)) {
        ptr = PyBytes_AS_STRING(it);
        len = PyBytes_GET_SIZE(it);
    } else if (PyByteArray_Check(it)) {
        ptr = PyByteArray_AS_STRING(it);
        len = PyByteArray_GET_SIZE(it);
    } else {
        return false;
    }
    return true;
}

// Resolver callbacks of bsq_*_stream_items (include/bsq.h).  The calling thread keeps the GIL for the whole
// call, like the reference does (src/tokenize.h:389-419 runs under it): nothing can resize a bytearray item or
// replace a list element while the pool threads read the items' headers and copy their bodies.
struct PyItems {
    PyObject **items;
    Py_ssize_t n;
};
// pool threads: bytes / bytearray items only (plain field reads); anything else is left to the caller (-1)
void py_resolve(void *vctx, int64_t lo, int64_t hi, const void **ptrs, int64_t *lens) {
    const PyItems &c = *static_cast<const PyItems *>(vctx);
    constexpr int64_t kAhead = 12;  // the walk is a chain of cache misses on object headers: keep a dozen in flight
    for (int64_t i = lo; i < std::min(hi, lo + kAhead); ++i) __builtin_prefetch(c.items[i]);
    for (int64_t i = lo; i < hi; ++i) {
        if (i + kAhead < hi) __builtin_prefetch(c.items[i + kAhead]);
        PyObject *it = c.items[i];
        if (PyBytes_Check(it)) {
            ptrs[i] = PyBytes_AS_STRING(it);
            lens[i] = PyBytes_GET_SIZE(it);
        } else if (PyByteArray_Check(it)) {
            ptrs[i] = PyByteArray_AS_STRING(it);
            lens[i] = PyByteArray_GET_SIZE(it);
        } else {
            lens[i] = -1;
        }
    }
}
// calling thread: str items (UTF-8 form via the C API); non-zero = a type the reference rejects, or a Python error
int py_fixup(void *vctx, int64_t i, const void **ptr, int64_t *len) {
    const PyItems &c = *static_cast<const PyItems *>(vctx);
    PyObject *it = c.items[i];
    if (PyUnicode_Check(it)) {
        Py_ssize_t size;
        const char *s = PyUnicode_AsUTF8AndSize(it, &size);
        if (s == nullptr) return 2;  // Python error set
        *ptr = s;
        *len = size;
        return 0;
    }
    if (PyBytes_Check(it)) {
        *ptr = PyBytes_AS_STRING(it);
        *len = PyBytes_GET_SIZE(it);
        return 0;
    }
    if (PyByteArray_Check(it)) {
        *ptr = PyByteArray_AS_STRING(it);
        *len = PyByteArray_GET_SIZE(it);
        return 0;
    }
    return 1;
}

void unpack_items(const py::sequence &batch, Unpacked &u, int /*nthreads*/ = 1) {
    PyObject *fast = PySequence_Fast(batch.ptr(), "batch must be a sequence");
    if (fast == nullptr) throw py::error_already_set();
    u.keepalive = py::reinterpret_steal<py::object>(fast);
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    u.ptrs.resize(static_cast<size_t>(n));
    u.lens.resize(static_cast<size_t>(n));
    const char *bad = "item was none of string, bytes, or numpy array of 8-bit integers. ";
    for (Py_ssize_t i = 0; i < n; ++i)
        if (!unpack_one(items[i], u.ptrs[i], u.lens[i])) throw py::value_error(bad);
}

// A host or device array argument of the *_packed entry points.
struct ArrayArg {
    py::object keepalive;
    const void *ptr = nullptr;
    int64_t count = 0;
    bool on_device = false;
    int device = -1;
};

ArrayArg array_arg(const py::object &obj, const char *what, int itemsize, const char *np_dtype, const py::object &torch_dtype) {
    ArrayArg a;
    Torch &t = Torch::get();
    if (py::isinstance(obj, t.mod.attr("Tensor"))) {
        if (!py::object(obj.attr("dtype")).equal(torch_dtype)) throw py::value_error(std::string(what) + ": wrong dtype");
        py::object c = obj.attr("contiguous")();
        a.keepalive = c;
        a.ptr = data_ptr(c);
        a.count = c.attr("numel")().cast<int64_t>();
        a.on_device = c.attr("is_cuda").cast<bool>();
        if (a.on_device) a.device = c.attr("get_device")().cast<int>();
        return a;
    }
    py::array arr = py::array::ensure(obj, py::array::c_style);
    if (!arr) throw py::value_error(std::string(what) + ": expected a numpy array or torch tensor");
    if (arr.itemsize() != itemsize || !arr.dtype().equal(py::dtype(np_dtype)))
        throw py::value_error(std::string(what) + ": wrong dtype, expected " + np_dtype);
    a.keepalive = arr;
    a.ptr = arr.data();
    a.count = arr.size();
    return a;
}

// ------------------------------------------------------------------ FlatFile (src/fxstats.cpp:24-200)
// Same constructor overloads, methods and properties as the reference's class.  Sequences come
// back as bytearray like there.  Additive: `pinned=` (hold the file in page-locked memory so
// batches DMA straight from it), `packed()` (zero-copy numpy views of residues + offsets) and
// being accepted directly by Tokenizer.batch_tokenize / batch_onehot_encode (no Python objects
// per sequence).
class FlatFile {
public:
    // prefault: 0 / False plain mapping, 1 / True populated at open, 2 populated AND page-locked in place (direct DMA)
    static int mode_of(bool pinned, int prefault) {
        return pinned ? BSQ_FF_PINNED : (prefault >= 2 ? BSQ_FF_MMAP_REGISTERED : (prefault ? BSQ_FF_MMAP_PREFAULT : BSQ_FF_MMAP));
    }
    FlatFile(const std::string &path, py::ssize_t maxseqlen, bool pinned, int prefault) : path_(path) {
        check(bsq_flatfile_open(&f_, path.c_str(), maxseqlen, mode_of(pinned, prefault)));
    }
    FlatFile(const std::string &inpath, const std::string &outpath, bool pinned, int prefault)
        : path_(outpath.empty() ? inpath + ".ff" : outpath) {
        int64_t n = 0, longest = 0;
        check(bsq_flatfile_make(inpath.c_str(), outpath.c_str(), &n, &longest));
        check(bsq_flatfile_open(&f_, path_.c_str(), longest, mode_of(pinned, prefault)));
    }
    FlatFile(const FlatFile &) = delete;
    FlatFile &operator=(const FlatFile &) = delete;
    ~FlatFile() { bsq_flatfile_close(f_); }

    const bsq_flatfile *handle() const { return f_; }
    const std::string &path() const { return path_; }
    int64_t nseqs() const { return bsq_flatfile_nseqs(f_); }
    int64_t seq_offset() const { return bsq_flatfile_seq_offset(f_); }
    int64_t max_seq_len() const { return bsq_flatfile_max_seq_len(f_); }
    bool is_pinned() const { return bsq_flatfile_is_pinned(f_) != 0; }

    py::bytearray access(int64_t i) const {
        if (i < 0 || i >= nseqs()) throw py::index_error("Accessing sequence out of range");  // src/fxstats.cpp:129
        const int64_t *o = bsq_flatfile_offsets(f_);
        return py::bytearray(reinterpret_cast<const char *>(bsq_flatfile_bytes(f_) + o[i]), static_cast<size_t>(o[i + 1] - o[i]));
    }
    py::list range_access(py::ssize_t i, py::ssize_t j, py::ssize_t step) const {  // src/fxstats.cpp:121-126
        if (step == 0) throw py::value_error("step must be nonzero");
        py::list ret;
        for (py::ssize_t idx = i; step > 0 ? idx < j : idx > j; idx += step) ret.append(access(idx));
        return ret;
    }
    py::list slice_access(const py::slice &slc) const {  // src/fxstats.cpp:106-114
        size_t start = 0, stop = 0, step = 0, len = 0;
        if (!slc.compute(static_cast<size_t>(nseqs()), &start, &stop, &step, &len)) throw py::error_already_set();
        return range_access(static_cast<py::ssize_t>(start), static_cast<py::ssize_t>(stop), static_cast<py::ssize_t>(step));
    }
    // The reference indexes its index array twice (`access(ptr[ptr[i]])`, src/fxstats.cpp:95-105);
    // here entry i selects sequence idx[i].
    py::list array_access(const py::array &idx) const {
        py::array_t<uint64_t, py::array::forcecast> ids(idx);
        py::list ret;
        const uint64_t *p = ids.data();
        for (py::ssize_t i = 0; i < ids.size(); ++i) ret.append(access(static_cast<int64_t>(p[i])));
        return ret;
    }
    py::object getitem(py::ssize_t idx) const {  // src/fxstats.cpp:176-187
        if (idx < 0) {
            if (idx < -static_cast<py::ssize_t>(nseqs())) throw py::index_error("For a negative index, idx must be >= -len(x)");
            idx += nseqs();
        }
        return access(idx);
    }
    py::array indptr() const {  // a copy, like src/fxstats.cpp:115-120
        py::array_t<uint64_t> ret(static_cast<py::ssize_t>(nseqs() + 1));
        std::memcpy(ret.mutable_data(), bsq_flatfile_offsets(f_), sizeof(uint64_t) * static_cast<size_t>(nseqs() + 1));
        return ret;
    }
    // zero-copy (residues uint8, offsets[start..stop] int64) views; they keep `self` alive
    py::tuple packed(py::ssize_t start, const py::object &stop_o, const py::object &self) const {
        const int64_t n = nseqs();
        int64_t stop = stop_o.is_none() ? n : stop_o.cast<int64_t>();
        if (start < 0 || stop < start || stop > n) throw py::index_error("Accessing sequence out of range");
        const int64_t *o = bsq_flatfile_offsets(f_);
        py::array_t<uint8_t> bytes({static_cast<py::ssize_t>(o[n])}, {1}, bsq_flatfile_bytes(f_), self);
        py::array_t<int64_t> offs({static_cast<py::ssize_t>(stop - start + 1)}, {8}, o + start, self);
        py::detail::array_proxy(bytes.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
        py::detail::array_proxy(offs.ptr())->flags &= ~py::detail::npy_api::NPY_ARRAY_WRITEABLE_;
        return py::make_tuple(bytes, offs);
    }

private:
    std::string path_;
    bsq_flatfile *f_ = nullptr;
};

// src/fxstats.cpp:136-151.  The reference's __next__ advances *before* the first item is
// read and yields the iterator itself, so `for it in ff` visits sequences 1..n-1 through
// `it.seq` / `it.sequence`; mirrored as is.  (`ff[i]`, `ff.access(...)`, `ff.packed()` see all.)
struct FlatFileIterator {
    const FlatFile *ff;
    int64_t start, stop;
    FlatFileIterator &next() {
        if (++start == stop) throw py::stop_iteration("End of iterator");
        return *this;
    }
    py::bytearray sequence() const { return ff->access(start); }
};

py::list getstats(const py::sequence &items) {  // src/fxstats.cpp:202-219
    py::list out;
    for (const auto &item : items) {
        const std::string path = item.cast<std::string>();
        int64_t *lens = nullptr, n = 0;
        check(bsq_fastx_lengths(path.c_str(), &lens, &n));
        py::array_t<uint64_t> arr(static_cast<py::ssize_t>(n));
        for (int64_t i = 0; i < n; ++i) arr.mutable_data()[i] = static_cast<uint64_t>(lens[i]);
        bsq_free(lens);
        out.append(arr);
    }
    return out;
}

// ------------------------------------------------------------------ the Tokenizer class
class Tokenizer {
public:
    Tokenizer(const std::string &key, bool eos, bool bos, bool padchar) : eos_(eos), bos_(bos), padchar_(padchar) {
        const int rc = bsq_tokenizer_init(&tok_, key.c_str(), eos, bos, padchar);
        if (rc != BSQ_OK) throw std::runtime_error(bsq_last_error());  // src/tokenize.h:78
    }

    // ---- introspection (src/tokenize.cpp:52-63, :99-106)
    std::string key() const { return tok_.key; }
    int nchars() const { return tok_.nchars; }
    size_t alphabet_size() const { return static_cast<size_t>(tok_.alphabet_size); }
    int bos() const { return tok_.bos_id; }
    int eos() const { return tok_.eos_id; }
    int pad() const { return tok_.pad_id; }
    bool is_padded() const { return padchar_; }
    bool includes_bos() const { return bos_; }
    bool includes_eos() const { return eos_; }

    std::vector<int32_t> ids() const {
        std::vector<int32_t> out;
        char buf[8];
        for (int32_t id = -128; id < tok_.nchars + 3; ++id)
            if (bsq_tokenizer_lookup(&tok_, id, buf, sizeof(buf)) > 0) out.push_back(id);
        return out;
    }
    py::dict lut() const {  // id -> decoded text
        py::dict d;
        char buf[8];
        for (int32_t id : ids()) {
            const int n = bsq_tokenizer_lookup(&tok_, id, buf, sizeof(buf));
            d[py::int_(id)] = py::reinterpret_steal<py::str>(PyUnicode_DecodeLatin1(buf, n, nullptr));
        }
        return d;
    }
    std::string token_map() const {  // "id:text;id:text" (reference order is unordered_map order)
        std::string s;
        char buf[8];
        for (int32_t id : ids()) {
            const int n = bsq_tokenizer_lookup(&tok_, id, buf, sizeof(buf));
            if (!s.empty()) s += ';';
            s += std::to_string(id) + ':' + std::string(buf, static_cast<size_t>(n));
        }
        return s;
    }
    py::dict token_decoder() const {  // id -> every byte that maps to it (src/tokenize.h:65-71)
        std::map<int, std::string> sets;
        for (int b = 0; b < 256; ++b) sets[tok_.lut[b]] += static_cast<char>(b);
        py::dict d;
        for (const auto &kv : sets) d[py::int_(kv.first)] = py::bytes(kv.second);
        return d;
    }
    py::tuple getstate() const { return py::make_tuple(key(), eos_, bos_, padchar_); }

    // ---- batch_tokenize (src/tokenize.cpp:82-98)
    py::object batch_tokenize(const py::sequence &batch, py::ssize_t padlen, const std::string &destchar, bool batch_first,
                              int nthreads, const py::object &device) const {
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        if (py::isinstance<FlatFile>(batch))  // additive: a whole FlatFile, no per-sequence objects
            return run_flatfile(batch.cast<const FlatFile &>(), 0, py::none(), py::none(), padlen, dc, kind, false, batch_first, device);
        if (padlen <= 0) throw py::value_error("batch tokenize requires padlen is provded.");
        return run_items(batch, padlen, dc, kind, /*onehot=*/false, batch_first, host_threads(nthreads), device);
    }

    // ---- batch_onehot_encode (src/tokenize.cpp:65-81); always (padlen, batch, alphabet_size)
    py::object batch_onehot_encode(const py::sequence &batch, py::ssize_t padlen, const std::string &destchar, int nthreads,
                                   const py::object &mask, const py::object &device) const {
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        if (py::isinstance<FlatFile>(batch))
            return run_flatfile(batch.cast<const FlatFile &>(), 0, py::none(), mask, padlen, dc, kind, true, false, device);
        if (padlen <= 0) throw py::value_error("batch tokenize requires padlen is provded.");
        nthreads = host_threads(nthreads);
        if (!py::isinstance<py::list>(mask)) return run_items(batch, padlen, dc, kind, /*onehot=*/true, false, nthreads, device);
        Unpacked u;
        unpack_items(batch, u, nthreads);
        // mask: a list with one uint8 array per sequence; entries that are not arrays mean
        // "no mask for this sequence" (getmaskptr, src/tokenize.h:372-380).
        std::vector<uint8_t> maskbuf;
        const uint8_t *maskptr = nullptr;
        if (py::isinstance<py::list>(mask)) {
            const py::list ml = mask.cast<py::list>();
            if (ml.size() < u.lens.size()) throw py::index_error("mask list is shorter than the batch");
            int64_t total = 0;
            for (int64_t l : u.lens) total += l;
            maskbuf.assign(static_cast<size_t>(total) + 32, 1);
            int64_t pos = 0;
            for (size_t i = 0; i < u.lens.size(); ++i) {
                py::object m = ml[i];
                if (py::isinstance<py::array>(m)) {
                    py::array_t<uint8_t, py::array::forcecast | py::array::c_style> arr(m);
                    if (arr.size() < u.lens[i]) throw py::value_error("mask entry shorter than its sequence");
                    std::memcpy(maskbuf.data() + pos, arr.data(), static_cast<size_t>(u.lens[i]));
                }
                pos += u.lens[i];
            }
            maskptr = maskbuf.data();
        }
        return run_host(u.ptrs.data(), u.lens.data(), static_cast<int64_t>(u.lens.size()), maskptr, padlen, dc, kind,
                        /*onehot=*/true, false, nthreads, device);
    }

    // ---- additive: already-packed input (host numpy / torch CPU, or torch CUDA = device resident)
    py::object tokenize_packed(const py::object &bytes, const py::object &offsets, py::ssize_t padlen,
                               const std::string &destchar, bool batch_first, const py::object &device, bool check_len) const {
        return run_packed(bytes, offsets, py::none(), padlen, destchar, false, batch_first, device, check_len);
    }
    py::object onehot_packed(const py::object &bytes, const py::object &offsets, py::ssize_t padlen, const std::string &destchar,
                             const py::object &mask, const py::object &device, bool check_len) const {
        return run_packed(bytes, offsets, mask, padlen, destchar, true, false, device, check_len);
    }

    // ---- additive: one packed host batch -> one shard per device (one process, one host thread per
    // GPU).  Shards are contiguous sequence ranges holding equal shares of the residues; shard g lives on devices[g]
    // as that device's own (n_g, padlen) / (padlen, n_g) tensor -- the per-device batches a DataParallel model
    // consumes (training/cnnpretrain.py:86).  Returns (list of tensors, bounds).
    py::tuple tokenize_sharded(const py::object &bytes, const py::object &offsets, py::ssize_t padlen, const std::string &destchar,
                               bool batch_first, const py::object &devices) const {
        Torch &t = Torch::get();
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        if (padlen <= 0) throw py::value_error("batch tokenize requires padlen is provded.");
        ArrayArg b = array_arg(bytes, "bytes", 1, "uint8", t.uint8);
        ArrayArg o = array_arg(offsets, "offsets", 8, "int64", t.int64);
        if (b.on_device || o.on_device) throw py::value_error("batch_tokenize_sharded takes host arrays (numpy / torch CPU, pinned or pageable)");
        if (o.count < 1) throw py::value_error("offsets needs at least one entry");
        const int64_t n = o.count - 1;
        const int64_t *ho = static_cast<const int64_t *>(o.ptr);
        if (n > 0 && (ho[0] < 0 || ho[n] > b.count)) throw py::value_error("offsets run past the end of bytes");
        check(bsq_check_lengths_host(ho, n, padlen, &tok_));
        std::vector<int> devs;
        if (devices.is_none()) {
            resolve_device(py::none());
            const int cnt = t.cuda.attr("device_count")().cast<int>();
            for (int g = 0; g < cnt; ++g) devs.push_back(g);
        } else {
            for (const auto &d : devices) devs.push_back(resolve_device(py::reinterpret_borrow<py::object>(d)));
        }
        if (devs.empty()) throw py::value_error("no devices given");
        const int G = static_cast<int>(devs.size());
        std::vector<int64_t> bounds(static_cast<size_t>(G) + 1);
        check(bsq_shard_bounds(ho, n, G, bounds.data()));
        py::list outs;
        std::vector<void *> optrs(G), streams(G);
        std::vector<bsq_stager *> stagers(G);
        std::vector<DeviceCtx *> ctxs(G);
        for (int g = 0; g < G; ++g) {
            const int64_t ng = bounds[g + 1] - bounds[g];
            std::vector<int64_t> shape = batch_first ? std::vector<int64_t>{ng, padlen} : std::vector<int64_t>{padlen, ng};
            py::object out = new_tensor(shape, t.dtype_of(dc, kind), devs[g]);
            optrs[g] = ng > 0 ? data_ptr(out) : nullptr;
            streams[g] = current_stream(devs[g]);
            ctxs[g] = &device_ctx(devs[g]);
            stagers[g] = ctxs[g]->stager;
            outs.append(out);
        }
        int rc;
        {
            py::gil_scoped_release nogil;
            std::vector<DeviceCtx *> order(ctxs);  // lock each device's context once, in address order
            std::sort(order.begin(), order.end());
            order.erase(std::unique(order.begin(), order.end()), order.end());
            if (order.size() != ctxs.size()) {
                rc = BSQ_ERR_ARG;
            } else {
                for (DeviceCtx *c : order) c->mu.lock();
                rc = bsq_tokenize_host_sharded(stagers.data(), streams.data(), G, static_cast<const uint8_t *>(b.ptr), ho, n, padlen, &tok_, 0,
                                               batch_first ? 1 : 0, kind, optrs.data(), bounds.data());
                for (DeviceCtx *c : order) c->mu.unlock();
            }
        }
        if (rc == BSQ_ERR_ARG && std::string(bsq_last_error()).empty()) throw py::value_error("devices must be distinct");
        check(rc);
        return py::make_tuple(outs, py::cast(bounds));
    }

    // ---- additive: sequences [start, stop) of a FlatFile; padlen <= 0 means the file's longest
    // sequence + bos + eos (what FlatFileDataset uses, bioseq/loaders.py:44)
    py::object tokenize_flatfile(const FlatFile &ff, py::ssize_t start, const py::object &stop, py::ssize_t padlen,
                                 const std::string &destchar, bool batch_first, const py::object &device) const {
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        return run_flatfile(ff, start, stop, py::none(), padlen, dc, kind, false, batch_first, device);
    }
    py::object onehot_flatfile(const FlatFile &ff, py::ssize_t start, const py::object &stop, py::ssize_t padlen,
                               const std::string &destchar, const py::object &mask, const py::object &device) const {
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        return run_flatfile(ff, start, stop, mask, padlen, dc, kind, true, false, device);
    }

    // ---- single-sequence onehot_encode (src/tokenize.cpp:8-48, src/tokenize.h:188-216)
    py::object onehot_encode(const py::object &seq, py::ssize_t padlen, const py::object &destchar, const py::object &device) const {
        // the reference registers three overloads whose default destchar differs: "f" for str
        // and bytearray, "B" for bytes (src/tokenize.cpp:31,39,48)
        const std::string dt = destchar.is_none() ? (PyBytes_Check(seq.ptr()) ? "B" : "f") : destchar.cast<std::string>();
        const void *ptr;
        int64_t len;
        Py_ssize_t size;
        if (PyUnicode_Check(seq.ptr())) {
            ptr = PyUnicode_AsUTF8AndSize(seq.ptr(), &size);
            if (ptr == nullptr) throw py::error_already_set();
            len = size;
        } else if (PyBytes_Check(seq.ptr())) {
            ptr = PyBytes_AS_STRING(seq.ptr());
            len = PyBytes_GET_SIZE(seq.ptr());
        } else if (PyByteArray_Check(seq.ptr())) {
            ptr = PyByteArray_AS_STRING(seq.ptr());
            len = PyByteArray_GET_SIZE(seq.ptr());
        } else {
            throw py::type_error("onehot_encode(): expected str, bytes or bytearray");
        }
        char dc;
        switch (dt.empty() ? 0 : (dt[0] & 223)) {  // src/tokenize.cpp:10-16: B H I F D only
            case 'B': dc = 'B'; break;
            case 'H': dc = 'h'; break;
            case 'I': dc = 'i'; break;
            case 'F': dc = 'f'; break;
            case 'D': dc = 'd'; break;
            default: throw py::value_error("Unsupported dtype: " + dt);
        }
        if (padlen > 0 && len > padlen) throw std::runtime_error("padlen is too short to accommodate sequence\n");
        // rows = max(len, padlen) + bos + eos (src/tokenize.h:195); pad rows only up to padlen (:210-214)
        const int64_t rows = std::max<int64_t>(len, padlen) + bos_ + eos_;
        if (rows == 0) return new_tensor({0, tok_.alphabet_size}, Torch::get().dtype_of(dc, bsq_kind_of_destchar(dc)), resolve_device(device));
        py::object out = run_host(&ptr, &len, 1, nullptr, rows, dc, bsq_kind_of_destchar(dc), true, false, 1, device);
        out = out.attr("reshape")(rows, tok_.alphabet_size);
        // the reference pads rows [len + bos + eos, padlen) only (src/tokenize.h:210-214); the kernel padded up to `rows`
        const int64_t zero_from = std::max<int64_t>(padlen, len + bos_ + eos_);
        if (padchar_ && zero_from < rows) out[py::slice(zero_from, rows, 1)].attr("zero_")();
        return out;
    }

    // ---- decode_tokens (src/tokenize.cpp:49-51, src/tokenize.h:131-183)
    py::object decode_tokens(const py::object &array, const py::object &device) const {
        Torch &t = Torch::get();
        py::object tens;  // device tensor viewed as raw bytes
        std::vector<int64_t> shape, strides;
        int itemsize;
        if (py::isinstance(array, t.mod.attr("Tensor"))) {
            py::object src = array;
            itemsize = src.attr("element_size")().cast<int>();
            shape = src.attr("shape").cast<std::vector<int64_t>>();
            strides = src.attr("stride")().cast<std::vector<int64_t>>();
            for (auto &s : strides) s *= itemsize;
            if (!src.attr("is_cuda").cast<bool>()) {
                const int dev = resolve_device(device);
                src = src.attr("contiguous")().attr("to")(t.device("cuda", dev));
                strides.assign(shape.size(), itemsize);
                for (int d = static_cast<int>(shape.size()) - 2; d >= 0; --d) strides[d] = strides[d + 1] * shape[d + 1];
            }
            tens = src;
        } else {
            py::array arr = py::array::ensure(array);
            if (!arr) throw py::type_error("decode_tokens(): expected a numpy array or torch tensor");
            itemsize = static_cast<int>(arr.itemsize());
            for (py::ssize_t d = 0; d < arr.ndim(); ++d) shape.push_back(arr.shape(d));
            const int nd = static_cast<int>(shape.size());
            if (nd > 2 || nd == 0) throw py::value_error("Currently supported: 1 or 2 dimensions for decoding tokens.");
            py::array flat = py::module_::import("numpy").attr("ascontiguousarray")(arr).attr("reshape")(-1).attr("view")("uint8");
            const int dev = resolve_device(device);
            tens = t.mod.attr("from_numpy")(flat.attr("copy")()).attr("to")(t.device("cuda", dev));
            strides.assign(shape.size(), itemsize);
            for (int d = nd - 2; d >= 0; --d) strides[d] = strides[d + 1] * shape[d + 1];
        }
        const int nd = static_cast<int>(shape.size());
        if (nd > 2 || nd == 0) throw py::value_error("Currently supported: 1 or 2 dimensions for decoding tokens.");
        const int64_t rows = nd == 1 ? 1 : shape[0], cols = nd == 1 ? shape[0] : shape[1];
        const int64_t rs = nd == 1 ? 0 : strides[0], cs = nd == 1 ? strides[0] : strides[1];
        const int dev = tens.attr("get_device")().cast<int>();
        void *st = current_stream(dev);
        py::object d_offs = new_tensor({rows + 1}, t.int64, dev);
        py::object d_tail = new_tensor({rows}, t.int32, dev);  // scratch: where each row's trailing <PAD> run begins
        int64_t total = 0;
        // Both passes behind one another with one synchronisation (bsq_decode_text): the text buffer is sized by a guess
        // -- 3 characters per token covers batches that are up to half <PAD> (5 each) -- and only a text that does
        // not fit costs a second, exactly sized pass 2.
        const int64_t guess = std::max<int64_t>(rows * cols * 3, 4096);
        py::object d_chars = new_tensor({guess}, t.uint8, dev);
        check(bsq_decode_text(dev, st, data_ptr(tens), itemsize, rows, cols, rs, cs, &tok_, static_cast<int64_t *>(data_ptr(d_offs)),
                              static_cast<int32_t *>(data_ptr(d_tail)), static_cast<uint8_t *>(data_ptr(d_chars)), guess, &total));
        if (total > guess) {
            d_chars = new_tensor({total}, t.uint8, dev);
            check(bsq_decode_chars(dev, st, data_ptr(tens), itemsize, rows, cols, rs, cs, &tok_,
                                   static_cast<const int64_t *>(data_ptr(d_offs)), static_cast<const int32_t *>(data_ptr(d_tail)),
                                   static_cast<uint8_t *>(data_ptr(d_chars))));
        }
        py::array_t<int64_t> h_offs = d_offs.attr("cpu")().attr("numpy")();
        const int64_t *o = h_offs.data();
        // The string objects are created first (sizes are known from the offsets); their bodies are then
        // filled straight from the device -> host ring by pool threads, without the GIL.  Every decoded
        // character is < 0x80 (ids of bytes >= 0x80 are rejected), so the strings are ASCII.
        py::list out(static_cast<size_t>(rows));
        std::vector<void *> dst(static_cast<size_t>(rows));
        for (int64_t r = 0; r < rows; ++r) {
            PyObject *s = PyUnicode_New(o[r + 1] - o[r], 127);
            if (s == nullptr) throw py::error_already_set();
            PyList_SET_ITEM(out.ptr(), r, s);
            dst[static_cast<size_t>(r)] = PyUnicode_1BYTE_DATA(s);
        }
        {
            DeviceCtx &ctx = device_ctx(dev);
            const uint8_t *chars_ptr = total > 0 ? static_cast<const uint8_t *>(data_ptr(d_chars)) : nullptr;
            const int nt = static_cast<int>(std::min<py::ssize_t>(get_num_threads(), 16));
            int rc;
            {
                py::gil_scoped_release nogil;
                std::lock_guard<std::mutex> g(ctx.mu);
                rc = bsq_fetch_rows(ctx.stager, st, chars_ptr, o, rows, dst.data(), nt);
            }
            check(rc);
        }
        if (nd == 1) return py::reinterpret_borrow<py::object>(PyList_GET_ITEM(out.ptr(), 0));
        return out;
    }

private:
    // The reference's calling convention, streamed: walk -> pinned pack -> copy -> kernel in one pass over the items
    // (bsq_*_stream_items).  The GIL stays with this thread for the whole call, as in the reference.
    py::object run_items(const py::sequence &batch, int64_t padlen, char dc, int kind, bool onehot, bool batch_first, int nthreads,
                         const py::object &device) const {
        PyObject *fast = PySequence_Fast(batch.ptr(), "batch must be a sequence");
        if (fast == nullptr) throw py::error_already_set();
        py::object keepalive = py::reinterpret_steal<py::object>(fast);
        PyItems items{PySequence_Fast_ITEMS(fast), PySequence_Fast_GET_SIZE(fast)};
        const int64_t n = items.n;
        // item types are checked while streaming; small batches are checked up front as well, so that their errors
        // precede any device use like the reference's (src/tokenize.h:389-419 unpacks before it allocates)
        if (n <= 4096)
            for (int64_t i = 0; i < n; ++i)
                if (!PyUnicode_Check(items.items[i]) && !PyBytes_Check(items.items[i]) && !PyByteArray_Check(items.items[i]))
                    throw py::value_error("item was none of string, bytes, or numpy array of 8-bit integers. ");
        const int dev = resolve_device(device);
        std::vector<int64_t> shape;
        if (onehot) shape = {padlen, n, tok_.alphabet_size};
        else if (batch_first) shape = {n, padlen};
        else shape = {padlen, n};
        py::object out = new_tensor(shape, Torch::get().dtype_of(dc, kind), dev);
        if (n == 0) return out;
        void *st = current_stream(dev);
        void *optr = data_ptr(out);
        DeviceCtx &ctx = device_ctx(dev);
        int rc;
        {
            std::lock_guard<std::mutex> g(ctx.mu);
            const int nt = nthreads > 0 ? nthreads : 1;
            if (onehot) rc = bsq_onehot_stream_items(ctx.stager, st, n, py_resolve, py_fixup, &items, padlen, &tok_, kind, optr, nt);
            else rc = bsq_tokenize_stream_items(ctx.stager, st, n, py_resolve, py_fixup, &items, padlen, &tok_, batch_first, kind, optr, nt);
        }
        if (rc != BSQ_OK && PyErr_Occurred()) throw py::error_already_set();  // raised inside py_fixup
        check(rc, onehot);
        return out;
    }

    // pack (pinned) -> stage -> launch.  ptrs/lens describe n host sequences.
    py::object run_host(const void *const *ptrs, const int64_t *lens, int64_t n, const uint8_t *mask, int64_t padlen, char dc,
                        int kind, bool onehot, bool batch_first, int nthreads, const py::object &device) const {
        const int dev = resolve_device(device);
        const int64_t extra = bos_ + eos_;
        for (int64_t i = 0; i < n; ++i)
            if (lens[i] + extra > padlen) {
                const std::string msg = "seq len + bos + eos > padlen: " + std::to_string(lens[i] + extra) + ", vs padlen " +
                                        std::to_string(padlen);
                if (onehot) throw py::value_error(msg);
                throw std::runtime_error(msg);
            }
        std::vector<int64_t> shape;
        if (onehot) shape = {padlen, n, tok_.alphabet_size};
        else if (batch_first) shape = {n, padlen};
        else shape = {padlen, n};
        py::object out = new_tensor(shape, Torch::get().dtype_of(dc, kind), dev);
        if (n == 0) return out;
        void *st = current_stream(dev);
        void *optr = data_ptr(out);
        DeviceCtx &ctx = device_ctx(dev);
        int rc;
        {
            // the GIL stays with this thread: ptrs borrow the items' buffers, which only it keeps unchanged
            std::lock_guard<std::mutex> g(ctx.mu);
            const int nt = nthreads > 0 ? nthreads : 1;
            if (mask == nullptr) {
                // gather -> copy -> kernel pipelined range by range (the pinned pack of the previous call is
                // waited for inside)
                if (onehot) rc = bsq_onehot_items(ctx.stager, ctx.pack, st, ptrs, lens, n, padlen, &tok_, kind, optr, nt);
                else rc = bsq_tokenize_items(ctx.stager, ctx.pack, st, ptrs, lens, n, padlen, &tok_, batch_first, kind, optr, nt);
            } else {
                rc = bsq_stager_sync_copies(ctx.stager);  // the pinned pack buffer may still be in flight
                if (rc == BSQ_OK) rc = bsq_pack_gather(ctx.pack, ptrs, lens, n, nt);
                if (rc == BSQ_OK)
                    rc = bsq_onehot_host(ctx.stager, st, bsq_pack_bytes(ctx.pack), bsq_pack_offsets(ctx.pack), mask, n, padlen,
                                         &tok_, kind, optr);
                if (rc == BSQ_OK) rc = bsq_stager_sync_copies(ctx.stager);  // mask is a local buffer
            }
        }
        check(rc, onehot);
        return out;
    }

    // FlatFile range -> device.  The file's offset table and residues are the packed form already.
    // mask (one-hot only): a uint8 array covering the whole file's residues, indexed like them.
    py::object run_flatfile(const FlatFile &ff, py::ssize_t start, const py::object &stop_o, const py::object &mask,
                            py::ssize_t padlen, char dc, int kind, bool onehot, bool batch_first, const py::object &device) const {
        Torch &t = Torch::get();
        const int64_t total = ff.nseqs();
        const int64_t stop = stop_o.is_none() ? total : stop_o.cast<int64_t>();
        if (start < 0 || stop < start || stop > total) throw py::index_error("Accessing sequence out of range");
        if (padlen <= 0) padlen = ff.max_seq_len() + bos_ + eos_;
        if (padlen <= 0) throw py::value_error("batch tokenize requires padlen is provded.");
        const int64_t n = stop - start;
        const int64_t *ho = bsq_flatfile_offsets(ff.handle()) + start;
        const uint8_t *hb = bsq_flatfile_bytes(ff.handle());
        ArrayArg m;
        if (!mask.is_none()) {
            m = array_arg(mask, "mask", 1, "uint8", t.uint8);
            if (m.on_device) throw py::value_error("mask must be a host array for a FlatFile batch");
            if (m.count < bsq_flatfile_offsets(ff.handle())[total]) throw py::value_error("mask shorter than the file's residues");
        }
        check(bsq_check_lengths_host(ho, n, padlen, &tok_), onehot);
        const int dev = resolve_device(device);
        std::vector<int64_t> shape;
        if (onehot) shape = {padlen, n, tok_.alphabet_size};
        else if (batch_first) shape = {n, padlen};
        else shape = {padlen, n};
        py::object out = new_tensor(shape, t.dtype_of(dc, kind), dev);
        if (n == 0) return out;
        void *st = current_stream(dev);
        void *optr = data_ptr(out);
        DeviceCtx &ctx = device_ctx(dev);
        int rc;
        {
            py::gil_scoped_release nogil;
            std::lock_guard<std::mutex> g(ctx.mu);
            if (onehot)
                rc = bsq_onehot_host(ctx.stager, st, hb, ho, static_cast<const uint8_t *>(m.ptr), n, padlen, &tok_, kind, optr);
            else
                rc = bsq_tokenize_host(ctx.stager, st, hb, ho, n, padlen, &tok_, batch_first, kind, optr);
            // a pinned file (or mask) is read asynchronously; it must outlive the copies
            if (rc == BSQ_OK && (ff.is_pinned() || m.ptr != nullptr)) rc = bsq_stager_sync_copies(ctx.stager);
        }
        check(rc, onehot);
        return out;
    }

    py::object run_packed(const py::object &bytes, const py::object &offsets, const py::object &mask, py::ssize_t padlen,
                          const std::string &destchar, bool onehot, bool batch_first, const py::object &device,
                          bool check_len) const {
        Torch &t = Torch::get();
        const char dc = destchar.empty() ? '\0' : destchar[0];
        const int kind = bsq_kind_of_destchar(dc);
        if (kind < 0) raise_status(kind);
        if (padlen <= 0) throw py::value_error("batch tokenize requires padlen is provded.");
        ArrayArg b = array_arg(bytes, "bytes", 1, "uint8", t.uint8);
        ArrayArg o = array_arg(offsets, "offsets", 8, "int64", t.int64);
        ArrayArg m;
        if (!mask.is_none()) m = array_arg(mask, "mask", 1, "uint8", t.uint8);
        if (o.count < 1) throw py::value_error("offsets needs at least one entry");
        if (b.on_device != o.on_device || (!mask.is_none() && m.on_device != b.on_device))
            throw py::value_error("bytes, offsets and mask must live on the same side (all host or all CUDA)");
        const int64_t n = o.count - 1;
        const int dev = b.on_device ? b.device : resolve_device(device);
        std::vector<int64_t> shape;
        if (onehot) shape = {padlen, n, tok_.alphabet_size};
        else if (batch_first) shape = {n, padlen};
        else shape = {padlen, n};
        void *st = current_stream(dev);
        int rc;
        py::object out;
        if (b.on_device) {
            // (also: lengths >= 0, offsets[0] >= 0, offsets[n] <= bytes.numel() -- the kernels index bytes with them)
            if (check_len) check(bsq_check_offsets_device(dev, st, static_cast<const int64_t *>(o.ptr), n, b.count, padlen, &tok_), onehot);
            out = new_tensor(shape, t.dtype_of(dc, kind), dev);
            if (onehot)
                rc = bsq_onehot(dev, st, static_cast<const uint8_t *>(b.ptr), static_cast<const int64_t *>(o.ptr),
                                static_cast<const uint8_t *>(m.ptr), n, padlen, &tok_, kind, data_ptr(out));
            else
                rc = bsq_tokenize(dev, st, static_cast<const uint8_t *>(b.ptr), static_cast<const int64_t *>(o.ptr), n, padlen,
                                  &tok_, batch_first, kind, data_ptr(out));
        } else {
            const int64_t *ho = static_cast<const int64_t *>(o.ptr);
            if (n > 0 && ho[0] < 0) throw py::value_error("offsets must start at or after 0");
            if (n > 0 && ho[n] > b.count) throw py::value_error("offsets run past the end of bytes");
            if (!mask.is_none() && m.count < b.count) throw py::value_error("mask shorter than bytes");
            check(bsq_check_lengths_host(ho, n, padlen, &tok_), onehot);
            out = new_tensor(shape, t.dtype_of(dc, kind), dev);
            DeviceCtx &ctx = device_ctx(dev);
            void *optr = data_ptr(out);
            py::gil_scoped_release nogil;
            std::lock_guard<std::mutex> g(ctx.mu);
            if (onehot)
                rc = bsq_onehot_host(ctx.stager, st, static_cast<const uint8_t *>(b.ptr), ho, static_cast<const uint8_t *>(m.ptr), n,
                                     padlen, &tok_, kind, optr);
            else
                rc = bsq_tokenize_host(ctx.stager, st, static_cast<const uint8_t *>(b.ptr), ho, n, padlen, &tok_, batch_first, kind,
                                       optr);
            // pageable arrays have been fully consumed (bounced through the pinned ring) on return;
            // pinned arrays are read asynchronously like any cudaMemcpyAsync source: the caller must
            // not overwrite them before the stream reaches this point.
        }
        check(rc, onehot);
        return out;
    }

    bsq_tokenizer tok_;
    bool eos_, bos_, padchar_;
};

}  // namespace bsqpy

PYBIND11_MODULE(cbioseq, m) {
    using bsqpy::FlatFile;
    using bsqpy::FlatFileIterator;
    using bsqpy::Tokenizer;
    m.doc() = "B200-native drop-in for bioseq's cbioseq tokenizer module (GPU batch tokenisation)";
    m.attr("__bsq_abi_version__") = bsq_abi_version();

    py::class_<FlatFileIterator>(m, "FlatFileIterator")
        .def(py::init<FlatFileIterator>())
        .def("__iter__", [](const FlatFileIterator &x) { return x; })
        .def("__next__", [](FlatFileIterator &x) { return x.next(); })
        .def_property_readonly("sequence", &FlatFileIterator::sequence)
        .def_property_readonly("seq", &FlatFileIterator::sequence);

    py::class_<FlatFile>(m, "FlatFile")
        .def(py::init<std::string, py::ssize_t, bool, bool>(), py::arg("inputfile"), py::arg("maxseqlen") = -1, py::kw_only(),
             py::arg("pinned") = false, py::arg("prefault") = 0)
        .def(py::init<std::string, std::string, bool, bool>(), py::arg("inputfile"), py::arg("outputfile"), py::kw_only(),
             py::arg("pinned") = false, py::arg("prefault") = 0)
        .def_property_readonly("path", &FlatFile::path)
        .def("access", &FlatFile::access)
        .def("access", &FlatFile::slice_access)
        .def("access", &FlatFile::range_access, py::arg("start"), py::arg("stop"), py::arg("step") = 1)
        .def("__len__", &FlatFile::nseqs)
        .def("nseqs", &FlatFile::nseqs)
        .def("size", &FlatFile::nseqs)
        .def("seq_offset", &FlatFile::seq_offset)
        .def("indptr", &FlatFile::indptr)
        .def_property_readonly("maxseqlen", &FlatFile::max_seq_len)
        .def_property_readonly("max_seq_len", &FlatFile::max_seq_len)
        .def_property_readonly("pinned", &FlatFile::is_pinned)
        .def("__iter__", [](const FlatFile &x) { return FlatFileIterator{&x, 0, x.nseqs()}; }, py::keep_alive<0, 1>())
        .def("__getitem__", &FlatFile::getitem)
        .def("__getitem__", &FlatFile::slice_access)
        .def("__getitem__", &FlatFile::array_access)
        .def("packed", [](py::object self, py::ssize_t start, const py::object &stop) {
                 return self.cast<const FlatFile &>().packed(start, stop, self);
             }, py::arg("start") = 0, py::arg("stop") = py::none());
    m.def("getstats", &bsqpy::getstats);

    py::class_<Tokenizer>(m, "Tokenizer")
        .def(py::init<std::string, bool, bool, bool>(), py::arg("key"), py::arg("eos") = false, py::arg("bos") = false,
             py::arg("padchar") = false)
        .def("onehot_encode", &Tokenizer::onehot_encode, py::arg("str"), py::arg("padlen") = 0, py::arg("destchar") = py::none(),
             py::arg("device") = py::none())
        .def("decode_tokens", &Tokenizer::decode_tokens, py::arg("tokenizer"), py::arg("device") = py::none())
        .def("lut", &Tokenizer::lut)
        .def("token_map", &Tokenizer::token_map)
        .def("token_decoder", &Tokenizer::token_decoder)
        .def("nchars", &Tokenizer::nchars)
        .def("batch_onehot_encode", &Tokenizer::batch_onehot_encode, py::arg("batch"), py::arg("padlen") = -1,
             py::arg("destchar") = "B", py::arg("nthreads") = 1, py::arg("mask") = py::none(), py::arg("device") = py::none())
        .def("batch_tokenize", &Tokenizer::batch_tokenize, py::arg("batch"), py::arg("padlen") = -1, py::arg("destchar") = "B",
             py::arg("batch_first") = false, py::arg("nthreads") = 1, py::arg("device") = py::none())
        .def("batch_tokenize_packed", &Tokenizer::tokenize_packed, py::arg("bytes"), py::arg("offsets"), py::arg("padlen") = -1,
             py::arg("destchar") = "B", py::arg("batch_first") = false, py::arg("device") = py::none(),
             py::arg("check_lengths") = true)
        .def("batch_onehot_encode_packed", &Tokenizer::onehot_packed, py::arg("bytes"), py::arg("offsets"), py::arg("padlen") = -1,
             py::arg("destchar") = "B", py::arg("mask") = py::none(), py::arg("device") = py::none(),
             py::arg("check_lengths") = true)
        .def("batch_tokenize_sharded", &Tokenizer::tokenize_sharded, py::arg("bytes"), py::arg("offsets"), py::arg("padlen") = -1,
             py::arg("destchar") = "B", py::arg("batch_first") = false, py::arg("devices") = py::none())
        .def("batch_tokenize_flatfile", &Tokenizer::tokenize_flatfile, py::arg("flatfile"), py::arg("start") = 0,
             py::arg("stop") = py::none(), py::arg("padlen") = -1, py::arg("destchar") = "B", py::arg("batch_first") = false,
             py::arg("device") = py::none())
        .def("batch_onehot_encode_flatfile", &Tokenizer::onehot_flatfile, py::arg("flatfile"), py::arg("start") = 0,
             py::arg("stop") = py::none(), py::arg("padlen") = -1, py::arg("destchar") = "B", py::arg("mask") = py::none(),
             py::arg("device") = py::none())
        .def("alphabet_size", &Tokenizer::alphabet_size)
        .def("bos", &Tokenizer::bos)
        .def("eos", &Tokenizer::eos)
        .def("pad", &Tokenizer::pad)
        .def_property_readonly("key", &Tokenizer::key)
        .def("is_padded", &Tokenizer::is_padded)
        .def("includes_bos", &Tokenizer::includes_bos)
        .def("includes_eos", &Tokenizer::includes_eos)
        .def(py::pickle([](const Tokenizer &t) { return t.getstate(); },
                        [](py::tuple s) {
                            return Tokenizer(s[0].cast<std::string>(), s[1].cast<bool>(), s[2].cast<bool>(), s[3].cast<bool>());
                        }));

    m.def("set_num_threads", &bsqpy::set_num_threads);
    m.def("get_num_threads", &bsqpy::get_num_threads);
    py::class_<bsqpy::Threading>(m, "Threading")
        .def(py::init<>())
        .def(py::init<py::ssize_t>())
        .def_property("nthreads", &bsqpy::Threading::get, &bsqpy::Threading::set)
        .def_property("p", &bsqpy::Threading::get, &bsqpy::Threading::set);
}
