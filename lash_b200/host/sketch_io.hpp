// On-disk formats of `lash sketch`:
//   {out}_sketches.bin     one zstd stream (level 3) of S::save records, list order   utils.rs:566-575
//   {out}_files.json       pretty JSON array of the input paths                        utils.rs:577-580
//   {out}_parameters.json  pretty JSON object of strings                               main.rs:249-276
// Record layouts (S::save / S::load, utils.rs:400-433 and :95-105, :202-222, :303-319):
//   HMH  hyperminhash 0.1.4 `serialize`: 16384 x u16 little-endian
//   ULL  ultraloglog 0.1.6 serde/bincode-1: {state: Vec<u8>}          = u64 LE length + bytes
//   HLL  streaming_algorithms 0.3.3 serde/bincode-1:
//        {alpha: f64, zero: usize, sum: f64, p: u8, m: Box<[u8]>}     = 8 + 8 + 8 + 1 + (8 + 2^p) bytes
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace lashhost {

size_t reg_bytes(int algo, int p);

// serialise n sketches (dense register arrays as lash_sketch_fetch returns them) into `path`
bool write_sketches(const std::string& path, int algo, int p, const void* regs, uint64_t n, int threads, std::string& err);
// read exactly n sketches; p_inout: expected precision (0 = take it from the first record)
bool read_sketches(const std::string& path, int algo, int* p_inout, uint64_t n, std::vector<uint8_t>& regs, std::string& err);

// serde_json::to_writer_pretty(&Vec<String>) / to_string_pretty(&Map) and their inverses
std::string json_pretty_string_array(const std::vector<std::string>& v);
std::string json_pretty_string_map(const std::map<std::string, std::string>& m);  // BTreeMap order == std::map order
bool json_parse_string_array(const std::string& text, std::vector<std::string>& out, std::string& err);
bool json_parse_string_map(const std::string& text, std::map<std::string, std::string>& out, std::string& err);

}  // namespace lashhost
