// internal.h -- shared between api_core.cpp (registry / parameter API) and dispatch.cpp
// (force dispatcher, SoA table builder, device staging).  Not installed.
#pragma once
#include <cstdint>
#include <cstddef>
#include <memory>
#include <mutex>
#include <algorithm>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "reboundx.h"
#include "rebx_b200.h"

namespace rebxh {

struct DeviceContext;   // dispatch.cpp

// Per-rebx_extras state that cannot live in the struct itself (Python allocates exactly
// seven pointers, reference reboundx/extras.py:327-333).
// A parameter whose interior value pointer was handed to the caller (rebx_get_param & co.): the reference's documented C
// idiom keeps such a pointer and writes through it between force evaluations (doc/c_quickstart.rst:49-54), which no API
// call announces.  The value's bits at the last check are remembered; every dispatch re-reads the (few) escaped values
// and rebuilds the tables when one moved.
struct Escaped { struct rebx_param* param; std::uint64_t bits; };

struct Side {
    std::uint64_t generation = 1;          // bumped whenever any parameter list may have changed
    std::vector<Escaped> escaped;          // see above; entries of freed parameters are nulled
    std::unordered_map<std::string, int> types;   // mirror of registered_params for O(1) rebx_get_type
    std::unique_ptr<DeviceContext, void (*)(DeviceContext*)> dev{nullptr, nullptr};
};

// Looks up (creating on demand) the side record of `rebx`.  Never returns NULL.
Side* side_of(struct rebx_extras* rebx);
// Drops the side record (device buffers included).  Safe to call twice.
void side_drop(struct rebx_extras* rebx);
// Marks every SoA table stale.
void mark_dirty(struct rebx_extras* rebx);
// Some list head was freed without a rebx pointer at hand (sim->free_particle_ap hook).
void mark_all_dirty();
// Re-reads the values whose pointers escaped through the getters; bumps the generation (tables rebuilt) if one changed.
// Called at the start of every dispatch / session call.  Cost: one 8-byte read per escaped parameter.
void refresh_escaped(Side* side);
std::uint64_t global_epoch();

// Interned parameter-name storage: every rebx_param created by this library points at one
// shared copy of its name, so the table builder can match names by pointer.
const char* intern_name(const char* name);

// List lookup without side effects (the exported rebx_get_param* additionally mark the
// tables dirty because the caller may write through the returned pointer).
struct rebx_param* find_param(struct rebx_node* ap, const char* name);
inline void* find_value(struct rebx_node* ap, const char* name) {
    struct rebx_param* p = find_param(ap, name);
    return p ? p->value : nullptr;
}

// Runs fn(thread, lo, hi) over [0, N) split into contiguous ranges on up to 32 host threads (one range on the
// calling thread when N is small).  Used by the table builder and the bulk parameter setters.
template <typename Fn>
inline void parallel_ranges(size_t N, Fn fn) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    size_t T = std::min<size_t>(std::min<size_t>(hw, 32), N/32768 + 1);
    if (T <= 1) { fn(0, size_t(0), N); return; }
    std::vector<std::thread> th;
    const size_t per = (N + T - 1)/T;
    for (size_t t = 0; t < T; t++) {
        const size_t lo = std::min(N, t*per), hi = std::min(N, lo + per);
        th.emplace_back([=] { fn((int)t, lo, hi); });
    }
    for (auto& x : th) x.join();
}

}  // namespace rebxh
