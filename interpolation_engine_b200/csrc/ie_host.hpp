// Host logic above the batch ABI: packing an inserts snapshot into the device table image and the
// JSON-level mirror of the reference's interp.rs functions (same names, argument meaning and
// error texts), every data-parallel step of which runs through the CUDA batch entry points.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ie_b200.h"

namespace ie_host {

// Builds the byte image of the device table (slots + key arena + value arena, ie_common.cuh).
bool build_table_image(uint64_t n, const uint8_t* keys, const uint64_t* key_offs, const uint8_t* vals, const uint64_t* val_offs,
                       const uint8_t* tags, const char* hhmm, const char* hhmmss, std::vector<uint8_t>* image, uint32_t* capacity,
                       std::string* why, bool compact = false, bool* any_balanced = nullptr);

ie_status_t call_json(ie_engine* e, const std::string& args_json, std::string* out_json, std::string* why);
void drop_engine(ie_engine* e);  // frees the snapshots registered through call_json for this engine

}  // namespace ie_host
