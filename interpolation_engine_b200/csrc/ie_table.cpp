// Packs an inserts snapshot (the `&Map<String, Value>` every interp.rs function takes) into the
// image of the device hash table: open addressing, linear probing, load factor <= 0.5, 64-byte
// slots with inline storage for keys / values of <= 16 bytes (ie_common.cuh).
#include <cstring>
#include <unordered_map>

#include "ie_common.cuh"
#include "ie_host.hpp"

namespace ie_host {

namespace {

struct Entry {
    const uint8_t* key; uint64_t key_len;
    const uint8_t* val; uint64_t val_len;
    uint32_t tag, index;
};

}  // namespace

bool build_table_image(uint64_t n, const uint8_t* keys, const uint64_t* key_offs, const uint8_t* vals, const uint64_t* val_offs,
                       const uint8_t* tags, const char* hhmm, const char* hhmmss, std::vector<uint8_t>* image, uint32_t* capacity,
                       std::string* why, bool compact, bool* any_balanced) {
    if (any_balanced) *any_balanced = false;
    if (n > 0x3FFFFFFFull) { *why = "too many inserts (max 2^30 - 1)"; return false; }
    std::vector<Entry> entries;
    entries.reserve(n + 2);
    for (uint64_t i = 0; i < n; ++i) {
        if (key_offs[i + 1] < key_offs[i] || val_offs[i + 1] < val_offs[i]) { *why = "offsets not monotone"; return false; }
        if (key_offs[i + 1] - key_offs[i] >= 0xFFFFFFFFull || val_offs[i + 1] - val_offs[i] > IE_VLEN_MAX) { *why = "key longer than 4 GiB or value longer than 32 MiB"; return false; }
        if (tags[i] > IE_TAG_OBJECT) { *why = "bad tag"; return false; }
        entries.push_back(Entry{keys + key_offs[i], key_offs[i + 1] - key_offs[i], vals + val_offs[i], val_offs[i + 1] - val_offs[i], tags[i], (uint32_t)i});
    }
    // interp.rs:96-104: the clock keys are answered before the map is consulted, so they shadow
    // inserts of the same name: appended last, and later duplicates win below.
    if (hhmm) entries.push_back(Entry{(const uint8_t*)"HH:MM", 5, (const uint8_t*)hhmm, std::strlen(hhmm), IE_TAG_STRING, (uint32_t)n});
    if (hhmmss) entries.push_back(Entry{(const uint8_t*)"HH:MM:SS", 8, (const uint8_t*)hhmmss, std::strlen(hhmmss), IE_TAG_STRING, (uint32_t)n + 1});

    // load factor <= 0.25 (<= 0.5 for very large maps): linear probe chains stay at 1-2 slots, and a
    // warp waits for its slowest lane
    uint64_t cap = 16;
    // (`compact`: <= 0.5 always — thousands of small snapshots packed side by side, ie_table_pack_many)
    const uint64_t want = (!compact && entries.size() <= (1ull << 22)) ? entries.size() * 4 : entries.size() * 2;
    while (cap < want) cap <<= 1;
    if (cap > (1ull << 31)) { *why = "table too large"; return false; }

    // arena sizes (16-byte aligned items)
    auto pad16 = [](uint64_t x) { return (x + 15) & ~uint64_t(15); };
    uint64_t key_bytes = 0, val_bytes = 0;
    for (auto& en : entries) {
        if (en.key_len > IE_INLINE_BYTES) key_bytes += pad16(en.key_len);
        if (en.val_len > IE_INLINE_BYTES) val_bytes += pad16(en.val_len);
    }
    const uint64_t slots_bytes = cap * sizeof(IeSlot);
    const uint64_t total = slots_bytes + key_bytes + val_bytes + 16;
    if ((total >> 4) > 0xFFFFFFFFull) { *why = "table image exceeds 64 GiB"; return false; }
    image->assign(total, 0);
    IeSlot* slots = reinterpret_cast<IeSlot*>(image->data());
    for (uint64_t i = 0; i < cap; ++i) slots[i].key_len = IE_SLOT_EMPTY;
    uint64_t kcur = slots_bytes, vcur = slots_bytes + key_bytes;
    const uint32_t mask = (uint32_t)(cap - 1);

    for (auto& en : entries) {
        const uint32_t h = ie_hash_bytes(en.key, (uint32_t)en.key_len);
        uint32_t idx = h & mask;
        IeSlot* s = nullptr;
        for (;;) {
            s = slots + idx;
            if (s->key_len == IE_SLOT_EMPTY) break;
            if (s->hash == h && s->key_len == en.key_len &&
                std::memcmp(image->data() + (uint64_t)s->key_off16 * 16, en.key, en.key_len) == 0) break;  // duplicate: overwrite
            idx = (idx + 1) & mask;
        }
        const bool fresh = s->key_len == IE_SLOT_EMPTY;
        if (fresh) {
            s->hash = h;
            s->key_len = (uint32_t)en.key_len;
            if (en.key_len <= IE_INLINE_BYTES) {
                std::memcpy(s->key_inline, en.key, en.key_len);
                s->key_off16 = (uint32_t)(((uint8_t*)s->key_inline - image->data()) >> 4);
            } else {
                std::memcpy(image->data() + kcur, en.key, en.key_len);
                s->key_off16 = (uint32_t)(kcur >> 4);
                kcur += pad16(en.key_len);
            }
        }
        const uint32_t vflags = ie_classify_value(en.val, en.val_len);
        if (any_balanced && (vflags & IE_VF_BALANCED)) *any_balanced = true;
        s->vl_tf = (uint32_t)en.val_len | (en.tag << 25) | (vflags << 28);
        s->entry = en.index;
        if (en.val_len <= IE_INLINE_BYTES) {
            std::memset(s->val_inline, 0, IE_INLINE_BYTES);
            std::memcpy(s->val_inline, en.val, en.val_len);
            s->val_off16 = (uint32_t)(((uint8_t*)s->val_inline - image->data()) >> 4);
        } else {
            std::memcpy(image->data() + vcur, en.val, en.val_len);
            s->val_off16 = (uint32_t)(vcur >> 4);
            vcur += pad16(en.val_len);
        }
    }
    *capacity = (uint32_t)cap;
    return true;
}

}  // namespace ie_host
