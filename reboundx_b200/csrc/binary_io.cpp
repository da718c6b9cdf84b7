// binary_io.cpp -- the REBOUNDx binary state format, written from scratch for this library
// (the data format either side of the force path).
//
// Format (reference src/output.c:33-89 describes it, src/reboundx.h:112-139,263-266 define the tags):
// a 64-byte text header, then nested {type:int32, pad, size:int64} fields.  A leaf field is followed by
// `size` payload bytes; a container field is followed by its children and closed by an END field, and its
// `size` spans children + END so that an unknown container can be skipped.
//
// Design here (not the reference's FILE* stream with ftell/fseek back-patching):
//   * writing: the tree is serialised into memory by a small builder that patches container sizes in
//     place.  All sizes are relative, so the per-particle records -- 99.9 % of a large file -- are built
//     as independent position-free segments by several threads and written out back to back;
//   * reading: the file is mapped and walked by a bounds-checked cursor (no per-field fread); parameter
//     names are interned once per distinct name and the registered type is cached, so loading 10^7
//     particle parameters costs one small allocation pair per parameter instead of four plus a
//     107-entry strcmp scan (reference src/input.c:64-104, src/core.c:1131-1138).
// The warning / error bits reported are the reference's (src/reboundx.h:147-165, src/input.c:629-690).
#include <algorithm>
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "internal.h"

using namespace rebxh;

namespace {

constexpr size_t kHeaderBytes = 64;
constexpr size_t kFieldBytes = sizeof(struct rebx_binary_field);
static_assert(kFieldBytes == 16, "rebx_binary_field must be {int32, pad, int64} like the reference's");

// ---- writer ---------------------------------------------------------------------------------------
class Builder {
public:
    std::vector<unsigned char> buf;
    void leaf(enum rebx_binary_field_type type, const void* data, size_t n) {
        put_field(type, (long)n);
        if (n) {
            const size_t at = buf.size();
            buf.resize(at + n);
            std::memcpy(&buf[at], data, n);
        }
    }
    size_t open(enum rebx_binary_field_type type) {       // returns the position of the container's header
        const size_t at = buf.size();
        put_field(type, 0);
        return at;
    }
    void close(size_t at) {                                // END field + size patch (children + END)
        put_field(REBX_BINARY_FIELD_TYPE_END, 0);
        const long size = (long)(buf.size() - (at + kFieldBytes));
        std::memcpy(&buf[at + offsetof(struct rebx_binary_field, size)], &size, sizeof size);
    }
    void append(const Builder& other) { buf.insert(buf.end(), other.buf.begin(), other.buf.end()); }
private:
    void put_field(enum rebx_binary_field_type type, long size) {
        struct rebx_binary_field f;
        std::memset(&f, 0, sizeof f);                      // padding bytes written as zero
        f.type = type; f.size = size;
        const size_t at = buf.size();
        buf.resize(at + sizeof f);
        std::memcpy(&buf[at], &f, sizeof f);
    }
};

// Payload bytes of a parameter value in the file (reference rebx_sizeof, src/core.c:1140-1177: STRING and
// every type added after the writer was written serialise as 0 bytes -- such a parameter does not survive
// a round trip in the reference either).
size_t payload_bytes(enum rebx_param_type type) {
    switch (type) {
        case REBX_TYPE_DOUBLE: return sizeof(double);
        case REBX_TYPE_INT:    return sizeof(int);
        case REBX_TYPE_VEC3D:  return sizeof(struct reb_vec3d);
        default:               return 0;
    }
}

void write_param(Builder& b, const struct rebx_param* param) {
    if (param->type == REBX_TYPE_POINTER || param->type == REBX_TYPE_ODE) return;       // src/output.c:137-139
    const size_t at = b.open(REBX_BINARY_FIELD_TYPE_PARAM);
    b.leaf(REBX_BINARY_FIELD_TYPE_PARAM_TYPE, &param->type, sizeof(param->type));
    b.leaf(REBX_BINARY_FIELD_TYPE_NAME, param->name, std::strlen(param->name) + 1);
    if (param->type == REBX_TYPE_FORCE) {           // stored as the force's name, re-linked on load (src/output.c:126-134)
        const struct rebx_force* force = static_cast<const struct rebx_force*>(param->value);
        b.leaf(REBX_BINARY_FIELD_TYPE_PARAM_VALUE, force->name, std::strlen(force->name) + 1);
    } else {
        b.leaf(REBX_BINARY_FIELD_TYPE_PARAM_VALUE, param->value, param->value ? payload_bytes(param->type) : 0);
    }
    b.close(at);
}

// Lists are push-front, so the file holds them tail first: reloading (push-front again) restores the order
// (reference rebx_write_list, src/output.c:212-257, re-walks the list per element; here it is reversed once).
template <typename Fn>
void for_each_tail_first(const struct rebx_node* head, Fn fn) {
    const struct rebx_node* small[32];                     // a particle carries a handful of parameters: no heap traffic
    size_t n = 0;
    const struct rebx_node* p = head;
    for (; p && n < 32; p = p->next) small[n++] = p;
    if (p) {                                               // long list (the registry): spill to the heap
        std::vector<const struct rebx_node*> nodes(small, small + n);
        for (; p; p = p->next) nodes.push_back(p);
        for (size_t i = nodes.size(); i-- > 0;) fn(nodes[i]->object);
        return;
    }
    for (size_t i = n; i-- > 0;) fn(small[i]->object);
}

void write_param_list(Builder& b, const struct rebx_node* ap) {
    const size_t at = b.open(REBX_BINARY_FIELD_TYPE_PARAM_LIST);
    for_each_tail_first(ap, [&](void* obj) { write_param(b, static_cast<const struct rebx_param*>(obj)); });
    b.close(at);
}

void write_particles(Builder& seg, const struct reb_particle* particles, size_t lo, size_t hi) {
    seg.buf.reserve((hi - lo)*192);                        // ~ one parameter per particle; grows geometrically beyond
    for (size_t i = lo; i < hi; i++) {
        const size_t at = seg.open(REBX_BINARY_FIELD_TYPE_PARTICLE);
        const int index = (int)i;
        seg.leaf(REBX_BINARY_FIELD_TYPE_PARTICLE_INDEX, &index, sizeof index);
        write_param_list(seg, static_cast<const struct rebx_node*>(particles[i].ap));
        seg.close(at);
    }
}

void write_structure(Builder& b, const struct rebx_extras* rebx) {
    const size_t at = b.open(REBX_BINARY_FIELD_TYPE_REBX_STRUCTURE);
    size_t l = b.open(REBX_BINARY_FIELD_TYPE_REGISTERED_PARAMETERS);
    for_each_tail_first(rebx->registered_params, [&](void* obj) {
        const struct rebx_param* p = static_cast<const struct rebx_param*>(obj);
        const size_t a = b.open(REBX_BINARY_FIELD_TYPE_REGISTERED_PARAM);
        b.leaf(REBX_BINARY_FIELD_TYPE_PARAM_TYPE, &p->type, sizeof(p->type));
        b.leaf(REBX_BINARY_FIELD_TYPE_NAME, p->name, std::strlen(p->name) + 1);
        b.close(a);
    });
    b.close(l);
    l = b.open(REBX_BINARY_FIELD_TYPE_ALLOCATED_FORCES);
    for_each_tail_first(rebx->allocated_forces, [&](void* obj) {
        const struct rebx_force* f = static_cast<const struct rebx_force*>(obj);
        const size_t a = b.open(REBX_BINARY_FIELD_TYPE_FORCE);
        b.leaf(REBX_BINARY_FIELD_TYPE_NAME, f->name, std::strlen(f->name) + 1);      // name first: the loader needs it to load the force
        write_param_list(b, f->ap);
        b.close(a);
    });
    b.close(l);
    // Operators are outside this library's scope: their lists are written empty.
    l = b.open(REBX_BINARY_FIELD_TYPE_ALLOCATED_OPERATORS); b.close(l);
    l = b.open(REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCES);
    for_each_tail_first(rebx->additional_forces, [&](void* obj) {
        const struct rebx_force* f = static_cast<const struct rebx_force*>(obj);
        const size_t a = b.open(REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCE);
        b.leaf(REBX_BINARY_FIELD_TYPE_NAME, f->name, std::strlen(f->name) + 1);
        b.close(a);
    });
    b.close(l);
    l = b.open(REBX_BINARY_FIELD_TYPE_PRE_TIMESTEP_MODIFICATIONS); b.close(l);
    l = b.open(REBX_BINARY_FIELD_TYPE_POST_TIMESTEP_MODIFICATIONS); b.close(l);
    b.close(at);
}

void make_header(char out[kHeaderBytes]) {
    // "REBOUNDx Binary File. Version: <version>\0<githash...>\0", 64 bytes (reference src/output.c:279-288)
    std::memset(out, 0, kHeaderBytes);
    const int len = snprintf(out, kHeaderBytes - 1, "REBOUNDx Binary File. Version: %s", rebx_version_str);
    if (len > 0 && (size_t)len + 1 < kHeaderBytes - 1) {
        const size_t room = kHeaderBytes - 1 - ((size_t)len + 1);
        std::strncpy(out + len + 1, rebx_githash_str, room);
    }
    out[kHeaderBytes - 1] = '\0';
}

// ---- reader ---------------------------------------------------------------------------------------
struct Cursor {
    const unsigned char* p;
    const unsigned char* end;
    bool field(struct rebx_binary_field& f) {
        if ((size_t)(end - p) < kFieldBytes) { p = end; return false; }
        std::memcpy(&f, p, kFieldBytes);
        p += kFieldBytes;
        return true;
    }
    bool has(long n) const { return n >= 0 && (size_t)(end - p) >= (size_t)n; }
    void skip(long n) { if (has(n)) p += n; else p = end; }
};

struct Loader {
    struct rebx_extras* rebx;
    int* warnings;
    std::unordered_map<std::string, const char*> names;      // distinct parameter names seen -> interned copy
    void warn(int bit) { *warnings |= bit; }

    const char* interned(const unsigned char* s, long n) {
        std::string key(reinterpret_cast<const char*>(s), strnlen(reinterpret_cast<const char*>(s), (size_t)n));
        auto it = names.find(key);
        if (it != names.end()) return it->second;
        const char* in = intern_name(key.c_str());
        names.emplace(std::move(key), in);
        return in;
    }

    // PARAM / REGISTERED_PARAM body up to its END.  Returns a param owned by the caller, or NULL.
    struct rebx_param* read_param(Cursor& c) {
        int type = REBX_TYPE_NONE;
        const char* name = nullptr;
        const unsigned char* value = nullptr; long vsize = -1;
        struct rebx_binary_field f;
        for (;;) {
            if (!c.field(f)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); break; }
            if (f.type == REBX_BINARY_FIELD_TYPE_END) break;
            if (!c.has(f.size)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); c.p = c.end; break; }
            switch (f.type) {
                case REBX_BINARY_FIELD_TYPE_PARAM_TYPE:
                    if (f.size == (long)sizeof(int)) std::memcpy(&type, c.p, sizeof(int)); else warn(REBX_INPUT_BINARY_ERROR_CORRUPT);
                    break;
                case REBX_BINARY_FIELD_TYPE_NAME:
                    if (f.size > 0) name = interned(c.p, f.size); else warn(REBX_INPUT_BINARY_ERROR_CORRUPT);
                    break;
                case REBX_BINARY_FIELD_TYPE_PARAM_VALUE:
                    value = c.p; vsize = f.size;
                    // a zero-byte payload (STRING, UINT32, ...) cannot be read back: the reference's fread of 0
                    // bytes fails and flags the file corrupt (src/input.c:46-58); same bits here
                    if (f.size == 0) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); value = nullptr; }
                    break;
                default:
                    warn(REBX_INPUT_BINARY_WARNING_FIELD_UNKNOWN);
                    break;
            }
            c.skip(f.size);
        }
        if (type == REBX_TYPE_NONE || name == nullptr) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return nullptr; }
        struct rebx_param* param = static_cast<struct rebx_param*>(malloc(sizeof *param));
        if (!param) { warn(REBX_INPUT_BINARY_ERROR_NO_MEMORY); return nullptr; }
        param->name = const_cast<char*>(name);
        param->type = static_cast<enum rebx_param_type>(type);
        param->value = nullptr;
        // a payload must have the size its type says (a FORCE reference: a NUL-terminated name), or later
        // readers of the value would run past it
        if (value && vsize > 0) {
            const size_t want = payload_bytes(param->type);
            const bool ok = param->type == REBX_TYPE_FORCE ? std::memchr(value, 0, (size_t)vsize) != nullptr
                                                           : (want == 0 || (size_t)vsize == want);
            if (!ok) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); value = nullptr; }
        }
        if (value && vsize > 0) {
            param->value = malloc((size_t)vsize);
            if (!param->value) { warn(REBX_INPUT_BINARY_ERROR_NO_MEMORY); free(param); return nullptr; }
            std::memcpy(param->value, value, (size_t)vsize);
        }
        return param;
    }

    static void drop(struct rebx_param* param) { if (param) { free(param->value); free(param); } }

    bool load_param(Cursor& c, struct rebx_node** ap) {
        struct rebx_param* param = read_param(c);
        if (!param) return false;
        if (param->value == nullptr) { warn(REBX_INPUT_BINARY_WARNING_PARAM_VALUE_NULL); drop(param); return false; }
        if (param->type == REBX_TYPE_FORCE) {       // value holds the force's name: link the re-created force
            char* fname = static_cast<char*>(param->value);
            struct rebx_force* force = rebx_get_force(rebx, fname);
            free(fname);
            param->value = nullptr;
            if (!force) { warn(REBX_INPUT_BINARY_WARNING_FORCE_PARAM_NOT_LOADED); drop(param); return false; }
            param->value = force;
        }
        struct rebx_node* node = static_cast<struct rebx_node*>(malloc(sizeof *node));
        if (!node) { warn(REBX_INPUT_BINARY_ERROR_NO_MEMORY); if (param->type != REBX_TYPE_FORCE) free(param->value); free(param); return false; }
        node->object = param; node->next = *ap; *ap = node;
        return true;
    }

    bool load_registered_param(Cursor& c) {
        struct rebx_param* param = read_param(c);
        if (!param) return false;
        const enum rebx_param_type type = param->type;
        const std::string name = param->name;
        drop(param);
        if (rebx_get_type(rebx, name.c_str()) != REBX_TYPE_NONE) return true;       // already known (loading into a live instance)
        rebx_register_param(rebx, name.c_str(), type);
        return rebx_get_type(rebx, name.c_str()) == type;
    }

    bool read_name(Cursor& c, std::string& out) {
        struct rebx_binary_field f;
        if (!c.field(f) || f.type != REBX_BINARY_FIELD_TYPE_NAME || f.size <= 0 || !c.has(f.size)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return false; }
        out.assign(reinterpret_cast<const char*>(c.p), strnlen(reinterpret_cast<const char*>(c.p), (size_t)f.size));
        c.skip(f.size);
        return true;
    }

    // Children of an object up to END: PARAM_LIST goes to `ap` (when given), anything else is skipped with a warning.
    bool read_object_tail(Cursor& c, struct rebx_node** ap) {
        struct rebx_binary_field f;
        for (;;) {
            if (!c.field(f)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return false; }
            if (f.type == REBX_BINARY_FIELD_TYPE_END) return true;
            if (f.type == REBX_BINARY_FIELD_TYPE_PARAM_LIST && ap) {
                if (!load_list(c, REBX_BINARY_FIELD_TYPE_PARAM, ap)) return false;
            } else {
                warn(REBX_INPUT_BINARY_WARNING_FIELD_UNKNOWN);
                c.skip(f.size);
            }
        }
    }

    bool load_force(Cursor& c) {
        std::string name;
        if (!read_name(c, name)) return false;
        struct rebx_force* force = rebx_load_force(rebx, name.c_str());
        if (!force) { warn(REBX_INPUT_BINARY_WARNING_FORCE_NOT_LOADED); return false; }
        return read_object_tail(c, &force->ap);
    }

    bool load_additional_force(Cursor& c) {
        std::string name;
        if (!read_name(c, name)) return false;
        struct rebx_force* force = rebx_get_force(rebx, name.c_str());      // re-created by the ALLOCATED_FORCES list
        if (!force) return false;
        if (!read_object_tail(c, nullptr)) return false;
        return rebx_add_force(rebx, force) != 0;
    }

    bool load_particle(Cursor& c) {
        struct rebx_binary_field f;
        int index = -1;
        if (!c.field(f) || f.type != REBX_BINARY_FIELD_TYPE_PARTICLE_INDEX || f.size != (long)sizeof(int) || !c.has(f.size)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return false; }
        std::memcpy(&index, c.p, sizeof index);
        c.skip(f.size);
        // the reference indexes sim->particles unchecked (src/input.c:363); a file for a larger ensemble is refused here
        if (index < 0 || (size_t)index >= rebx->sim->N) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return false; }
        return read_object_tail(c, reinterpret_cast<struct rebx_node**>(&rebx->sim->particles[index].ap));
    }

    // A list container's children up to its END; false only when the file is malformed (src/input.c:497-588).
    bool load_list(Cursor& c, enum rebx_binary_field_type expected, struct rebx_node** ap) {
        struct rebx_binary_field f;
        for (;;) {
            if (!c.field(f)) return false;
            if (f.type == REBX_BINARY_FIELD_TYPE_END) return true;
            if (f.type != expected) return false;
            Cursor item = {c.p, c.has(f.size) ? c.p + f.size : c.end};       // an item never reads past its own size
            bool ok = false; int bit = 0;
            switch (expected) {
                case REBX_BINARY_FIELD_TYPE_PARAM:            ok = load_param(item, ap);        bit = REBX_INPUT_BINARY_WARNING_PARAM_NOT_LOADED; break;
                case REBX_BINARY_FIELD_TYPE_REGISTERED_PARAM: ok = load_registered_param(item); bit = REBX_INPUT_BINARY_ERROR_REGISTERED_PARAM_NOT_LOADED; break;
                case REBX_BINARY_FIELD_TYPE_FORCE:            ok = load_force(item);            bit = REBX_INPUT_BINARY_WARNING_FORCE_NOT_LOADED; break;
                case REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCE: ok = load_additional_force(item); bit = REBX_INPUT_BINARY_WARNING_ADDITIONAL_FORCE_NOT_LOADED; break;
                case REBX_BINARY_FIELD_TYPE_PARTICLE:         ok = load_particle(item);         bit = REBX_INPUT_BINARY_WARNING_PARTICLE_PARAMS_NOT_LOADED; break;
                case REBX_BINARY_FIELD_TYPE_OPERATOR:         ok = false; bit = REBX_INPUT_BINARY_WARNING_OPERATOR_NOT_LOADED; break;   // operators: out of scope
                case REBX_BINARY_FIELD_TYPE_STEP:             ok = false; bit = REBX_INPUT_BINARY_WARNING_OPERATOR_NOT_LOADED | REBX_INPUT_BINARY_WARNING_STEP_NOT_LOADED; break;
                default:
                    rebx_error(rebx, "REBOUNDx Error. Reached default in rebx_load_list reading binary. Should never reach this case. Means we added a list to rebx and didn't add new case to load_list. Please report bug as Github issue.\n");
                    return false;
            }
            if (!ok) warn(bit);
            c.skip(f.size);
        }
    }

    // The PARTICLES container: one record per particle, each touching only its own particle's list.  A large file
    // whose records carry strictly increasing particle indices (every file this library or the reference writes) is
    // loaded by several threads, each with its own name cache and warning word; anything else goes record by record.
    bool load_particles(Cursor& c) {
        struct Item { const unsigned char* p; long size; };
        std::vector<Item> items;
        bool increasing = true, well_formed = true;
        long last_index = -1;
        {
            Cursor scan = c;
            struct rebx_binary_field f;
            for (;;) {
                if (!scan.field(f)) { well_formed = false; break; }
                if (f.type == REBX_BINARY_FIELD_TYPE_END) break;
                if (f.type != REBX_BINARY_FIELD_TYPE_PARTICLE || !scan.has(f.size)) { well_formed = false; break; }
                // peek at the record's PARTICLE_INDEX (its first field)
                struct rebx_binary_field g;
                int index = -1;
                if ((size_t)f.size >= kFieldBytes + sizeof(int)) {
                    std::memcpy(&g, scan.p, kFieldBytes);
                    if (g.type == REBX_BINARY_FIELD_TYPE_PARTICLE_INDEX && g.size == (long)sizeof(int)) std::memcpy(&index, scan.p + kFieldBytes, sizeof index);
                }
                if (index <= last_index) increasing = false;
                last_index = index;
                items.push_back({scan.p, f.size});
                scan.skip(f.size);
            }
        }
        if (!well_formed || !increasing || items.size() < 500000) return load_list(c, REBX_BINARY_FIELD_TYPE_PARTICLE, nullptr);
        std::vector<int> w(64, 0);
        parallel_ranges(items.size(), [&](int t, size_t lo, size_t hi) {
            Loader local{rebx, &w[(size_t)t], {}};
            for (size_t i = lo; i < hi; i++) {
                Cursor item = {items[i].p, items[i].p + items[i].size};
                if (!local.load_particle(item)) local.warn(REBX_INPUT_BINARY_WARNING_PARTICLE_PARAMS_NOT_LOADED);
            }
        });
        for (int bits : w) *warnings |= bits;
        return true;
    }

    void load_structure(Cursor& c) {
        struct rebx_binary_field f;
        for (;;) {
            if (!c.field(f)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return; }
            if (f.type == REBX_BINARY_FIELD_TYPE_END) return;
            Cursor body = {c.p, c.has(f.size) ? c.p + f.size : c.end};
            bool ok = true;
            switch (f.type) {
                case REBX_BINARY_FIELD_TYPE_REGISTERED_PARAMETERS:       ok = load_list(body, REBX_BINARY_FIELD_TYPE_REGISTERED_PARAM, &rebx->registered_params); break;
                case REBX_BINARY_FIELD_TYPE_ALLOCATED_FORCES:            ok = load_list(body, REBX_BINARY_FIELD_TYPE_FORCE, nullptr); break;
                case REBX_BINARY_FIELD_TYPE_ALLOCATED_OPERATORS:         ok = load_list(body, REBX_BINARY_FIELD_TYPE_OPERATOR, nullptr); break;
                case REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCES:           ok = load_list(body, REBX_BINARY_FIELD_TYPE_ADDITIONAL_FORCE, nullptr); break;
                case REBX_BINARY_FIELD_TYPE_PRE_TIMESTEP_MODIFICATIONS:  ok = load_list(body, REBX_BINARY_FIELD_TYPE_STEP, &rebx->pre_timestep_modifications); break;
                case REBX_BINARY_FIELD_TYPE_POST_TIMESTEP_MODIFICATIONS: ok = load_list(body, REBX_BINARY_FIELD_TYPE_STEP, &rebx->post_timestep_modifications); break;
                default: warn(REBX_INPUT_BINARY_WARNING_FIELD_UNKNOWN); break;
            }
            if (!ok) warn(REBX_INPUT_BINARY_ERROR_CORRUPT);
            c.skip(f.size);
        }
    }

    void load_snapshot(Cursor& c) {
        struct rebx_binary_field f;
        if (!c.field(f) || f.type != REBX_BINARY_FIELD_TYPE_SNAPSHOT) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return; }
        for (;;) {
            if (!c.field(f)) { warn(REBX_INPUT_BINARY_ERROR_CORRUPT); return; }
            if (f.type == REBX_BINARY_FIELD_TYPE_END) return;
            Cursor body = {c.p, c.has(f.size) ? c.p + f.size : c.end};
            switch (f.type) {
                case REBX_BINARY_FIELD_TYPE_REBX_STRUCTURE: load_structure(body); break;
                case REBX_BINARY_FIELD_TYPE_PARTICLES:
                    if (!load_particles(body)) warn(REBX_INPUT_BINARY_ERROR_CORRUPT);
                    break;
                default: warn(REBX_INPUT_BINARY_WARNING_LIST_UNKNOWN); break;
            }
            c.skip(f.size);
        }
    }
};

void check_header(const unsigned char* data, size_t n, int* warnings) {
    if (n < kHeaderBytes) { *warnings |= REBX_INPUT_BINARY_ERROR_CORRUPT; return; }
    char want[kHeaderBytes];
    make_header(want);
    // the version string (up to the first NUL) must match; the git hash behind it is ignored (src/input.c:608-612)
    if (std::strncmp(reinterpret_cast<const char*>(data), want, kHeaderBytes) != 0) *warnings |= REBX_INPUT_BINARY_WARNING_VERSION;
}

}  // namespace

extern "C" {

// reference src/output.c:272-293
void rebx_output_binary(struct rebx_extras* rebx, char* filename) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return; }
    FILE* of = fopen(filename, "wb");
    if (of == nullptr) { rebx_error(rebx, "REBOUNDx error: Can not open file passed to rebx_output_binary."); return; }
    if (rebx->pre_timestep_modifications || rebx->post_timestep_modifications || rebx->allocated_operators)
        reb_simulation_warning(rebx->sim, "REBOUNDx-B200: operators are outside this library's scope and were not written to the binary file.");
    const struct reb_simulation* sim = rebx->sim;
    const size_t N = sim->N;

    // particle records: independent position-free segments, built in parallel
    unsigned hw = std::thread::hardware_concurrency();
    const size_t T = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(hw ? hw : 1, 16), N/65536 + 1));
    std::vector<Builder> seg(T);
    {
        std::vector<std::thread> th;
        const size_t per = (N + T - 1)/T;
        for (size_t t = 0; t < T; t++) {
            const size_t lo = std::min(N, t*per), hi = std::min(N, lo + per);
            if (T == 1) write_particles(seg[t], sim->particles, lo, hi);
            else th.emplace_back([&seg, sim, t, lo, hi] { write_particles(seg[t], sim->particles, lo, hi); });
        }
        for (auto& x : th) x.join();
    }
    size_t particles_bytes = 0;
    for (const Builder& s : seg) particles_bytes += s.buf.size();

    Builder head;                                           // SNAPSHOT header, the rebx structure, PARTICLES header
    const size_t snap = head.open(REBX_BINARY_FIELD_TYPE_SNAPSHOT);
    write_structure(head, rebx);
    const size_t plist = head.open(REBX_BINARY_FIELD_TYPE_PARTICLES);
    Builder tail;                                           // END(PARTICLES), END(SNAPSHOT)
    const size_t e1 = tail.open(REBX_BINARY_FIELD_TYPE_END); (void)e1;
    const size_t e2 = tail.open(REBX_BINARY_FIELD_TYPE_END); (void)e2;
    const long plist_size = (long)(particles_bytes + kFieldBytes);
    const long snap_size = (long)(head.buf.size() - (snap + kFieldBytes) + particles_bytes + 2*kFieldBytes);
    std::memcpy(&head.buf[plist + offsetof(struct rebx_binary_field, size)], &plist_size, sizeof plist_size);
    std::memcpy(&head.buf[snap + offsetof(struct rebx_binary_field, size)], &snap_size, sizeof snap_size);

    char header[kHeaderBytes];
    make_header(header);
    bool ok = fwrite(header, 1, kHeaderBytes, of) == kHeaderBytes;
    ok = ok && fwrite(head.buf.data(), 1, head.buf.size(), of) == head.buf.size();
    for (const Builder& s : seg) ok = ok && (s.buf.empty() || fwrite(s.buf.data(), 1, s.buf.size(), of) == s.buf.size());
    ok = ok && fwrite(tail.buf.data(), 1, tail.buf.size(), of) == tail.buf.size();
    if (fclose(of) != 0) ok = false;
    if (!ok) rebx_error(rebx, "REBOUNDx error: writing the binary file failed (disk full?).");
}

// reference src/input.c:617-632
void rebx_init_extras_from_binary(struct rebx_extras* rebx, const char* const filename, enum rebx_input_binary_messages* warnings) {
    if (rebx->sim == nullptr) { rebx_error(rebx, ""); return; }
    int w = (int)*warnings;
    const int fd = open(filename, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        if (fd >= 0) close(fd);
        *warnings = static_cast<enum rebx_input_binary_messages>(w | REBX_INPUT_BINARY_ERROR_NOFILE);
        return;
    }
    const size_t n = (size_t)st.st_size;
    void* map = n ? mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (n && map == MAP_FAILED) { *warnings = static_cast<enum rebx_input_binary_messages>(w | REBX_INPUT_BINARY_ERROR_NO_MEMORY); return; }
    const unsigned char* data = static_cast<const unsigned char*>(map);
    check_header(data, n, &w);
    if (n >= kHeaderBytes) {
        Loader L{rebx, &w, {}};
        Cursor c = {data + kHeaderBytes, data + n};
        L.load_snapshot(c);
    }
    if (map) munmap(map, n);
    mark_dirty(rebx);
    *warnings = static_cast<enum rebx_input_binary_messages>(w);
}

// reference src/input.c:634-690
struct rebx_extras* rebx_create_extras_from_binary(struct reb_simulation* sim, const char* const filename) {
    if (sim == nullptr) {
        fprintf(stderr, "REBOUNDx Error: Simulation pointer passed to rebx_create_extras_from_binary was NULL.\n");
        return nullptr;
    }
    enum rebx_input_binary_messages warnings = REBX_INPUT_BINARY_WARNING_NONE;
    struct rebx_extras* rebx = static_cast<struct rebx_extras*>(malloc(sizeof *rebx));     // no default registrations: the file brings its own
    if (!rebx) return nullptr;
    rebx_initialize(sim, rebx);
    rebx_init_extras_from_binary(rebx, filename, &warnings);
    static const struct { int bit; bool error; const char* msg; } table[] = {
        {REBX_INPUT_BINARY_ERROR_NOFILE, true, "REBOUNDx: Cannot open binary file. Check filename."},
        {REBX_INPUT_BINARY_ERROR_CORRUPT, true, "REBOUNDx: Binary file is unreadable. Please open an issue on Github mentioning the version of REBOUND and REBOUNDx you are using and include the binary file."},
        {REBX_INPUT_BINARY_ERROR_NO_MEMORY, true, "REBOUNDx: Ran out of system memory."},
        {REBX_INPUT_BINARY_ERROR_REBX_NOT_LOADED, true, "REBOUNDx: REBOUNDx structure couldn't be loaded."},
        {REBX_INPUT_BINARY_ERROR_REGISTERED_PARAM_NOT_LOADED, true, "REBOUNDx: At least one registered parameter was not loaded. This typically indicates the binary is corrupt or was saved with an incompatible version to the current one being used."},
        {REBX_INPUT_BINARY_WARNING_PARAM_NOT_LOADED, false, "REBOUNDx: At least one force or operator parameter was not loaded from the binary file. This typically indicates the binary is corrupt or was saved with an incompatible version to the current one being used."},
        {REBX_INPUT_BINARY_WARNING_PARTICLE_PARAMS_NOT_LOADED, false, "REBOUNDx: At least one particle's parameters were not loaded from the binary file."},
        {REBX_INPUT_BINARY_WARNING_FORCE_NOT_LOADED, false, "REBOUNDx: At least one force was not loaded from the binary file. If binary was created with a newer version of REBOUNDx, a particular force may not be implemented in your current version of REBOUNDx."},
        {REBX_INPUT_BINARY_WARNING_OPERATOR_NOT_LOADED, false, "REBOUNDx: At least one operator was not loaded from the binary file. If binary was created with a newer version of REBOUNDx, a particular force may not be implemented in your current version of REBOUNDx."},
        {REBX_INPUT_BINARY_WARNING_STEP_NOT_LOADED, false, "REBOUNDx: At least one operator step was not loaded from the binary file."},
        {REBX_INPUT_BINARY_WARNING_ADDITIONAL_FORCE_NOT_LOADED, false, "REBOUNDx: At least one force was not added to the simulation. If binary was created with a newer version of REBOUNDx, a particular force may not be implemented in your current version of REBOUNDx."},
        {REBX_INPUT_BINARY_WARNING_FIELD_UNKNOWN, false, "REBOUNDx: Unknown field found in binary file. Any unknown fields not loaded.  This can happen if the binary was created with a later version of REBOUNDx than the one used to read it."},
        {REBX_INPUT_BINARY_WARNING_LIST_UNKNOWN, false, "REBOUNDx: Unknown list in the REBOUNDx structure wasn't loaded. This can happen if the binary was created with a later version of REBOUNDx than the one used to read it."},
        {REBX_INPUT_BINARY_WARNING_PARAM_VALUE_NULL, false, "REBOUNDx: The value of at least one parameter was not loaded. This can happen if a custom structure was added by the user as a parameter. See Parameters.ipynb jupyter notebook example."},
        {REBX_INPUT_BINARY_WARNING_VERSION, false, "REBOUNDx: Binary file was saved with a different version of REBOUNDx. Binary format might have changed. Check that effects and parameters are loaded as expected."},
        {REBX_INPUT_BINARY_WARNING_FORCE_PARAM_NOT_LOADED, false, "REBOUNDx: A force parameter failed to load from the list of REBOUNDx implemented forces. Custom forces can't be saved to a REBOUNDx binary, and function points must be reset when a simulation is reloaded."},
    };
    for (const auto& row : table)
        if ((int)warnings & row.bit) { if (row.error) reb_simulation_error(sim, row.msg); else reb_simulation_warning(sim, row.msg); }
    return rebx;
}

// reference src/input.c:692-700
FILE* rebx_input_inspect_binary(const char* const filename, enum rebx_input_binary_messages* warnings) {
    FILE* inf = fopen(filename, "rb");
    if (!inf) { *warnings = static_cast<enum rebx_input_binary_messages>((int)*warnings | REBX_INPUT_BINARY_ERROR_NOFILE); return nullptr; }
    unsigned char head[kHeaderBytes];
    const size_t got = fread(head, 1, kHeaderBytes, inf);
    int w = (int)*warnings;
    check_header(head, got, &w);
    *warnings = static_cast<enum rebx_input_binary_messages>(w);
    return inf;
}

// reference src/input.c:702-712
struct rebx_binary_field rebx_input_read_binary_field(FILE* inf) {
    struct rebx_binary_field field;
    if (fread(&field, sizeof field, 1, inf) == 1) return field;
    std::memset(&field, 0, sizeof field);
    return field;
}

// reference src/input.c:56-58
void rebx_input_skip_binary_field(FILE* inf, long field_size) { fseek(inf, field_size, SEEK_CUR); }

}  // extern "C"
