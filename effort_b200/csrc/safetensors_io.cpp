// safetensors_io.cpp -- host-side reader of the reference's bucketed-safetensors model directory.
//
// The on-disk contract is the reference's, unchanged (helpers/safetensors.swift):
//   <model>.safetensors.index.json   {"weight_map": {tensor name: file name}}                 :38-85, :105-110
//   <model>-%05d-of-%05d.safetensors  u64 little-endian header size, JSON header
//                                     {name: {"dtype", "shape", "data_offsets": [a, b]}, "__metadata__": {...}},
//                                     then the tensor bytes at 8 + header size + a                   :153-183
// TensorLoader semantics kept (:136-216): a name that is not in the index is retried as name + ".weight"; only
// BF16 / F16 / F32 are accepted; b - a must equal prod(shape) * sizeof(dtype).  Errors are return codes instead of
// precondition failures.  Files are mapped read-only and stay mapped until the loader is closed, so the returned
// pointers can be handed straight to cudaMemcpy (or wrapped by numpy) without another copy.
// No CUDA in this file: it is the "loader" row of SURVEY.md section 8(f), and it is unit-tested on the CPU.
#include <fcntl.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/effort_b200.h"

namespace {

// ---- a JSON subset parser: objects, arrays, strings (with escapes), numbers, true/false/null -------------
struct JVal {
    enum Kind { kNull, kBool, kNum, kStr, kArr, kObj } kind = kNull;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;  // insertion order = file order
    const JVal* get(const std::string& k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++;
    }
    bool lit(const char* s) {
        size_t n = strlen(s);
        if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    bool hex4(unsigned& v) {
        if (end - p < 4) return false;
        v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else return false;
        }
        return true;
    }
    bool string(std::string& out) {
        if (p >= end || *p != '"') return false;
        p++;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                p++;
                if (p >= end) return false;
                char c = *p++;
                switch (c) {
                    case '"': out += '"'; break;
                    case '\\': out += '\\'; break;
                    case '/': out += '/'; break;
                    case 'b': out += '\b'; break;
                    case 'f': out += '\f'; break;
                    case 'n': out += '\n'; break;
                    case 'r': out += '\r'; break;
                    case 't': out += '\t'; break;
                    case 'u': {
                        unsigned cp;
                        if (!hex4(cp)) return false;
                        if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                            p += 2;
                            unsigned lo;
                            if (!hex4(lo)) return false;
                            if (lo < 0xDC00 || lo >= 0xE000) return false;  // not a low surrogate
                            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        utf8(out, cp);
                        break;
                    }
                    default: return false;
                }
            } else {
                out += *p++;
            }
        }
        if (p >= end) return false;
        p++;
        return true;
    }
    JVal value(int depth = 0) {
        JVal v;
        ws();
        if (p >= end || depth > 64) { ok = false; return v; }
        if (*p == '{') {
            p++;
            v.kind = JVal::kObj;
            ws();
            if (p < end && *p == '}') { p++; return v; }
            while (ok) {
                ws();
                std::string k;
                if (!string(k)) { ok = false; break; }
                ws();
                if (p >= end || *p != ':') { ok = false; break; }
                p++;
                JVal c = value(depth + 1);
                v.obj.emplace_back(std::move(k), std::move(c));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == '}') { p++; break; }
                ok = false;
            }
        } else if (*p == '[') {
            p++;
            v.kind = JVal::kArr;
            ws();
            if (p < end && *p == ']') { p++; return v; }
            while (ok) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p < end && *p == ',') { p++; continue; }
                if (p < end && *p == ']') { p++; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v.kind = JVal::kStr;
            if (!string(v.str)) ok = false;
        } else if (lit("true")) { v.kind = JVal::kBool; v.b = true; }
        else if (lit("false")) { v.kind = JVal::kBool; }
        else if (lit("null")) { v.kind = JVal::kNull; }
        else {
            char* e = nullptr;
            std::string tmp(p, (size_t)(end - p) < 64 ? (size_t)(end - p) : 64);
            v.num = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) { ok = false; return v; }
            p += e - tmp.c_str();
            v.kind = JVal::kNum;
        }
        return v;
    }
};

bool parse_json(const char* data, size_t n, JVal& out) {
    JParser ps{data, data + n};
    out = ps.value();
    ps.ws();
    return ps.ok && ps.p == ps.end;
}

struct TensorRec {
    int dtype;  // EFFORT_ST_*
    std::vector<int64_t> shape;
    uint64_t begin, end;  // byte offsets inside the data section
};

struct StFile {
    int fd = -1;
    const unsigned char* map = nullptr;
    size_t size = 0;
    size_t data_off = 0;  // 8 + header size
    std::map<std::string, TensorRec> tensors;
    std::vector<std::string> order;
    std::string description;
    ~StFile() {
        if (map) munmap((void*)map, size);
        if (fd >= 0) close(fd);
    }
};

int dtype_code(const std::string& s) {
    if (s == "F16") return EFFORT_ST_F16;
    if (s == "BF16") return EFFORT_ST_BF16;
    if (s == "F32") return EFFORT_ST_F32;
    return -1;
}
size_t dtype_size(int code) { return code == EFFORT_ST_F32 ? 4 : 2; }

int open_file(const std::string& path, std::unique_ptr<StFile>& out) {
    std::unique_ptr<StFile> f(new StFile());
    f->fd = open(path.c_str(), O_RDONLY);
    if (f->fd < 0) return EFFORT_ENOTLOADED;
    struct stat st;
    if (fstat(f->fd, &st) != 0 || st.st_size < 8) return EFFORT_EINVAL;
    f->size = (size_t)st.st_size;
    void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
    if (m == MAP_FAILED) return EFFORT_ENOMEM;
    f->map = (const unsigned char*)m;
    uint64_t hsize = 0;
    for (int i = 7; i >= 0; i--) hsize = (hsize << 8) | f->map[i];  // little endian
    if (hsize > f->size - 8) return EFFORT_EINVAL;
    f->data_off = 8 + (size_t)hsize;
    JVal hdr;
    if (!parse_json((const char*)f->map + 8, (size_t)hsize, hdr) || hdr.kind != JVal::kObj) return EFFORT_EINVAL;
    for (auto& kv : hdr.obj) {
        if (kv.first == "__metadata__") {
            if (const JVal* d = kv.second.get("description"))
                if (d->kind == JVal::kStr) f->description = d->str;
            continue;
        }
        const JVal *dt = kv.second.get("dtype"), *sh = kv.second.get("shape"), *off = kv.second.get("data_offsets");
        if (!dt || !sh || !off || dt->kind != JVal::kStr || sh->kind != JVal::kArr || off->kind != JVal::kArr ||
            off->arr.size() != 2)
            return EFFORT_EINVAL;
        TensorRec r;
        r.dtype = dtype_code(dt->str);  // unsupported types are kept (-1) and refused when asked for
        uint64_t count = 1;
        const double kMaxExact = 9007199254740992.0;  // 2^53: beyond it a JSON number is not an exact integer
        for (auto& d : sh->arr) {
            if (d.kind != JVal::kNum || !(d.num >= 0) || d.num > kMaxExact || d.num != (double)(uint64_t)d.num) return EFFORT_EINVAL;
            r.shape.push_back((int64_t)d.num);
            if (d.num != 0 && count > (uint64_t)kMaxExact / (uint64_t)d.num) return EFFORT_EINVAL;  // overflow
            count *= (uint64_t)d.num;
        }
        for (int k = 0; k < 2; k++)
            if (off->arr[k].kind != JVal::kNum || !(off->arr[k].num >= 0) || off->arr[k].num > kMaxExact ||
                off->arr[k].num != (double)(uint64_t)off->arr[k].num)
                return EFFORT_EINVAL;
        r.begin = (uint64_t)off->arr[0].num;
        r.end = (uint64_t)off->arr[1].num;
        if (r.end < r.begin || r.end > f->size - f->data_off) return EFFORT_EINVAL;
        if (r.dtype >= 0 && r.end - r.begin != count * dtype_size(r.dtype)) return EFFORT_ESHAPE;  // safetensors.swift:182
        f->order.push_back(kv.first);
        f->tensors.emplace(kv.first, std::move(r));
    }
    out = std::move(f);
    return EFFORT_OK;
}

}  // namespace

struct effort_loader {
    std::string dir;
    std::map<std::string, std::string> weight_map;          // tensor -> file name
    std::vector<std::string> names;                         // index order
    std::map<std::string, std::unique_ptr<StFile>> files;   // opened lazily, kept mapped
};

static std::string join_path(const std::string& a, const std::string& b) {
    if (a.empty()) return b;
    return a.back() == '/' ? a + b : a + "/" + b;
}

extern "C" int effort_loader_open(const char* dir, const char* model, effort_loader_t** out) {
    if (!dir || !model || !out) return EFFORT_EINVAL;
    *out = nullptr;
    const std::string index_path = join_path(dir, std::string(model) + ".safetensors.index.json");
    int fd = open(index_path.c_str(), O_RDONLY);
    if (fd < 0) return EFFORT_ENOTLOADED;
    std::string text;
    char buf[65536];
    ssize_t n;
    while ((n = read(fd, buf, sizeof buf)) > 0) text.append(buf, (size_t)n);
    close(fd);
    JVal idx;
    if (!parse_json(text.data(), text.size(), idx) || idx.kind != JVal::kObj) return EFFORT_EINVAL;
    const JVal* wm = idx.get("weight_map");
    if (!wm || wm->kind != JVal::kObj) return EFFORT_EINVAL;
    std::unique_ptr<effort_loader> L(new effort_loader());
    L->dir = dir;
    for (auto& kv : wm->obj) {
        if (kv.second.kind != JVal::kStr) return EFFORT_EINVAL;
        if (L->weight_map.emplace(kv.first, kv.second.str).second) L->names.push_back(kv.first);
    }
    *out = L.release();
    return EFFORT_OK;
}

extern "C" void effort_loader_close(effort_loader_t* L) { delete L; }

extern "C" int effort_loader_count(const effort_loader_t* L) { return L ? (int)L->names.size() : 0; }

extern "C" const char* effort_loader_name(const effort_loader_t* L, int i) {
    if (!L || i < 0 || i >= (int)L->names.size()) return nullptr;
    return L->names[i].c_str();
}

// hasTensor (safetensors.swift:132-134) with fetchTensor's ".weight" fallback (:141-146)
static const std::string* resolve(const effort_loader* L, const char* name, std::string& key) {
    key = name;
    auto it = L->weight_map.find(key);
    if (it == L->weight_map.end()) {
        key += ".weight";
        it = L->weight_map.find(key);
        if (it == L->weight_map.end()) return nullptr;
    }
    return &it->second;
}

extern "C" int effort_loader_has(const effort_loader_t* L, const char* name) {
    if (!L || !name) return 0;
    std::string key;
    return resolve(L, name, key) != nullptr;
}

extern "C" int effort_loader_tensor(effort_loader_t* L, const char* name, effort_tensor_info_t* info) {
    if (!L || !name || !info) return EFFORT_EINVAL;
    memset(info, 0, sizeof *info);
    std::string key;
    const std::string* fname = resolve(L, name, key);
    if (!fname) return EFFORT_ENOTLOADED;  // "not found in the safetensors lib" (:148)
    auto it = L->files.find(*fname);
    if (it == L->files.end()) {
        std::unique_ptr<StFile> f;
        int rc = open_file(join_path(L->dir, *fname), f);
        if (rc) return rc;
        it = L->files.emplace(*fname, std::move(f)).first;
    }
    const StFile& f = *it->second;
    auto t = f.tensors.find(key);
    if (t == f.tensors.end()) return EFFORT_ENOTLOADED;
    const TensorRec& r = t->second;
    if (r.dtype < 0) return EFFORT_EINVAL;  // only BF16 / F16 / F32 (:176)
    if (r.shape.size() > EFFORT_ST_MAX_DIMS) return EFFORT_ESHAPE;
    info->dtype = r.dtype;
    info->ndim = (int)r.shape.size();
    for (size_t d = 0; d < r.shape.size(); d++) info->shape[d] = r.shape[d];
    info->data = f.map + f.data_off + r.begin;
    info->nbytes = (size_t)(r.end - r.begin);
    return EFFORT_OK;
}

// convertBF16 (model.swift / aux.metal: bfloat -> half): exact widening to fp32, then round-to-nearest-even to fp16
extern "C" int effort_bf16_to_f16(const uint16_t* src, uint16_t* dst, size_t n) {
    if (!src || !dst) return EFFORT_EINVAL;
    for (size_t i = 0; i < n; i++) {
        const uint32_t x = (uint32_t)src[i] << 16;
        const uint32_t sign = (x >> 16) & 0x8000u;
        const int32_t e = (int32_t)((x >> 23) & 0xFF) - 127;
        uint32_t m = x & 0x7FFFFFu;
        uint16_t h;
        if (((x >> 23) & 0xFF) == 0xFF) {
            h = (uint16_t)(sign | 0x7C00u | (m ? 0x200u : 0u));                       // inf / nan
        } else if (e > 15) {
            h = (uint16_t)(sign | 0x7C00u);                                           // overflow -> inf
        } else if (e >= -14) {                                                        // normal half
            uint32_t mant = m >> 13, rest = m & 0x1FFFu;
            uint32_t v = ((uint32_t)(e + 15) << 10) | mant;
            if (rest > 0x1000u || (rest == 0x1000u && (v & 1u))) v++;
            h = (uint16_t)(sign | v);
        } else if (e >= -25) {                                                        // subnormal half
            m |= 0x800000u;
            const int shift = -e - 14 + 13;  // bits dropped
            uint32_t mant = m >> shift, rest = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
            if (rest > half || (rest == half && (mant & 1u))) mant++;
            h = (uint16_t)(sign | mant);
        } else {
            h = (uint16_t)sign;
        }
        dst[i] = h;
    }
    return EFFORT_OK;
}


// ---------------------------------------------------------------------------------------------------------------
// writer: TensorSaver (helpers/safetensors.swift:38-85) + saveSafetensors (:222-280)
// ---------------------------------------------------------------------------------------------------------------
// files[i] collects named host tensors; save() writes "<model>-%05d-of-%05d.safetensors" (u64 little-endian header
// size, JSON header {name: {"dtype", "shape", "data_offsets"}, "__metadata__": {"description": ...}}, tensor bytes in
// header order) and "<model>.safetensors.index.json" {"weight_map": {name: file}}.  The saver copies nothing: the
// caller's buffers must stay valid until effort_saver_save returns.
namespace {
struct SaveRec {
    std::string name;
    int dtype;
    std::vector<int64_t> shape;
    const void* data;
    size_t nbytes;
};
void json_escape(std::string& o, const std::string& s) {
    o += '"';
    for (unsigned char c : s) {
        switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); o += b; }
                else o += (char)c;
        }
    }
    o += '"';
}
const char* dtype_name(int code) { return code == EFFORT_ST_F16 ? "F16" : code == EFFORT_ST_BF16 ? "BF16" : "F32"; }
bool write_all(int fd, const void* buf, size_t n) {
    const char* p = (const char*)buf;
    while (n) {
        ssize_t w = write(fd, p, n > ((size_t)1 << 30) ? ((size_t)1 << 30) : n);
        if (w <= 0) return false;
        p += w;
        n -= (size_t)w;
    }
    return true;
}
}  // namespace

struct effort_saver {
    std::string dir, model, description;
    std::vector<std::vector<SaveRec>> files;
};

extern "C" int effort_saver_open(const char* dir, const char* model, const char* description, effort_saver_t** out) {
    if (!dir || !model || !out) return EFFORT_EINVAL;
    effort_saver* s = new (std::nothrow) effort_saver();
    if (!s) return EFFORT_ENOMEM;
    s->dir = dir;
    s->model = model;
    s->description = description ? description : "Bucket weights format, see mixtral-kolinko at github";  // safetensors.swift:249
    *out = s;
    return EFFORT_OK;
}

extern "C" void effort_saver_close(effort_saver_t* s) { delete s; }

extern "C" int effort_saver_add(effort_saver_t* s, int file_index, const char* name, int dtype, int ndim,
                                const int64_t* shape, const void* data_host, size_t nbytes) {
    if (!s || !name || file_index < 0 || file_index > 100000 || ndim < 0 || ndim > EFFORT_ST_MAX_DIMS || (ndim && !shape) ||
        (nbytes && !data_host))
        return EFFORT_EINVAL;
    if (dtype != EFFORT_ST_F16 && dtype != EFFORT_ST_F32 && dtype != EFFORT_ST_BF16) return EFFORT_EINVAL;  // :231-247
    uint64_t count = 1;
    for (int d = 0; d < ndim; d++) {
        if (shape[d] < 0) return EFFORT_EINVAL;
        count *= (uint64_t)shape[d];
    }
    if (count * (uint64_t)dtype_size(dtype) != (uint64_t)nbytes) return EFFORT_ESHAPE;
    while ((int)s->files.size() <= file_index) s->files.emplace_back();  // subscript setter, :49-61
    for (auto& f : s->files)
        for (auto& r : f)
            if (r.name == name) return EFFORT_ESTATE;  // a tensor name maps to ONE file in the index
    SaveRec r;
    r.name = name; r.dtype = dtype; r.shape.assign(shape, shape + ndim); r.data = data_host; r.nbytes = nbytes;
    s->files[(size_t)file_index].push_back(std::move(r));
    return EFFORT_OK;
}

extern "C" int effort_saver_save(effort_saver_t* s) {
    if (!s) return EFFORT_EINVAL;
    mkdir(s->dir.c_str(), 0777);
    std::string index = "{\n  \"weight_map\": {";
    bool first = true;
    char fname[256];
    for (size_t id = 0; id < s->files.size(); id++) {
        snprintf(fname, sizeof(fname), "%s-%05d-of-%05d.safetensors", s->model.c_str(), (int)id + 1, (int)s->files.size());  // :67
        std::string hdr = "{";
        uint64_t off = 0;
        for (auto& r : s->files[id]) {
            json_escape(hdr, r.name);
            hdr += ":{\"dtype\":\"";
            hdr += dtype_name(r.dtype);
            hdr += "\",\"shape\":[";
            for (size_t d = 0; d < r.shape.size(); d++) { if (d) hdr += ','; hdr += std::to_string(r.shape[d]); }
            hdr += "],\"data_offsets\":[" + std::to_string(off) + "," + std::to_string(off + r.nbytes) + "]},";
            off += r.nbytes;
            index += first ? "\n    " : ",\n    ";
            first = false;
            json_escape(index, r.name);
            index += ": ";
            json_escape(index, fname);
        }
        hdr += "\"__metadata__\":{\"description\":";
        json_escape(hdr, s->description);
        hdr += "}}";
        while (hdr.size() % 8) hdr += ' ';  // keep the tensor bytes 8-byte aligned (allowed padding of the header)
        const std::string path = join_path(s->dir, fname);
        int fd = open(path.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) return EFFORT_EINVAL;
        uint64_t n = hdr.size();
        unsigned char le[8];
        for (int b = 0; b < 8; b++) le[b] = (unsigned char)(n >> (8 * b));
        bool ok = write_all(fd, le, 8) && write_all(fd, hdr.data(), hdr.size());
        for (auto& r : s->files[id]) ok = ok && write_all(fd, r.data, r.nbytes);
        ok = (close(fd) == 0) && ok;
        if (!ok) return EFFORT_EINVAL;
    }
    index += "\n  }\n}\n";
    const std::string ipath = join_path(s->dir, s->model + ".safetensors.index.json");
    int fd = open(ipath.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
    if (fd < 0) return EFFORT_EINVAL;
    bool ok = write_all(fd, index.data(), index.size());
    ok = (close(fd) == 0) && ok;
    return ok ? EFFORT_OK : EFFORT_EINVAL;
}
