// internal.h -- shared between api_core.cpp (registry / parameter API) and dispatch.cpp
// (force dispatcher, SoA table builder, device staging).  Not installed.
#pragma once
#include <cstdint>
#include <cstddef>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "reboundx.h"
#include "rebx_b200.h"

namespace rebxh {

struct DeviceContext;   // dispatch.cpp

// Per-rebx_extras state that cannot live in the struct itself (Python allocates exactly
// seven pointers, reference reboundx/extras.py:327-333).
struct Side {
    std::uint64_t generation = 1;          // bumped whenever any parameter list may have changed
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

}  // namespace rebxh
