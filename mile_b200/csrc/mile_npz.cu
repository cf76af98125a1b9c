// mile_npz.cu -- native writer of the reference's sample files (src/training/callbacks.py:17-44: one compressed .npz per
// chain and kept position, members 'fcn.layer0.bias', 'fcn.layer0.kernel', ...).  A reference-sized run writes 12 000 of
// them; through np.savez_compressed that is 0.7-1.0 ms per file under the GIL (threads do not help) and was the dominant
// cost of a whole run (tools/full_run.py: 11.7 s of which ~8 s were file writing).  Here a batch of files is deflated and
// written by a few host threads: plain zip archives (method 8), readable by np.load / the reference's
// load_samples_from_dir unchanged.  Host code only (no CUDA).
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mile_b200.h"

namespace {

void put16(std::vector<unsigned char>& b, uint32_t v) { b.push_back(v & 255); b.push_back((v >> 8) & 255); }
void put32(std::vector<unsigned char>& b, uint32_t v) { put16(b, v & 0xffff); put16(b, v >> 16); }

struct Member { std::string name; const unsigned char* header; int header_len; long n_floats; };

// one archive into `out` (cleared first); returns false on a zlib error
bool build_npz(const std::vector<Member>& members, const float* data, std::vector<unsigned char>& out,
               std::vector<unsigned char>& raw, std::vector<unsigned char>& comp) {
  out.clear();
  std::vector<unsigned char> cd;
  long off = 0;
  int n = 0;
  for (const Member& m : members) {
    const size_t usize = (size_t)m.header_len + (size_t)m.n_floats * 4;
    raw.resize(usize);
    memcpy(raw.data(), m.header, m.header_len);
    memcpy(raw.data() + m.header_len, data + off, (size_t)m.n_floats * 4);
    off += m.n_floats;
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data(), (uInt)usize);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    comp.resize(deflateBound(&zs, (uLong)usize));
    zs.next_in = raw.data(); zs.avail_in = (uInt)usize;
    zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
    const int rc = deflate(&zs, Z_FINISH);
    const size_t csize = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END) return false;
    const uint32_t lho = (uint32_t)out.size();
    // local file header
    put32(out, 0x04034b50u); put16(out, 20); put16(out, 0); put16(out, 8); put16(out, 0); put16(out, 0x21);   // time 0, date 1980-01-01
    put32(out, crc); put32(out, (uint32_t)csize); put32(out, (uint32_t)usize);
    put16(out, (uint32_t)m.name.size()); put16(out, 0);
    out.insert(out.end(), m.name.begin(), m.name.end());
    out.insert(out.end(), comp.begin(), comp.begin() + csize);
    // central directory entry
    put32(cd, 0x02014b50u); put16(cd, 20); put16(cd, 20); put16(cd, 0); put16(cd, 8); put16(cd, 0); put16(cd, 0x21);
    put32(cd, crc); put32(cd, (uint32_t)csize); put32(cd, (uint32_t)usize);
    put16(cd, (uint32_t)m.name.size()); put16(cd, 0); put16(cd, 0); put16(cd, 0); put16(cd, 0); put32(cd, 0);
    put32(cd, lho);
    cd.insert(cd.end(), m.name.begin(), m.name.end());
    ++n;
  }
  const uint32_t cdo = (uint32_t)out.size();
  out.insert(out.end(), cd.begin(), cd.end());
  put32(out, 0x06054b50u); put16(out, 0); put16(out, 0); put16(out, n); put16(out, n);
  put32(out, (uint32_t)cd.size()); put32(out, cdo); put16(out, 0);
  return true;
}

}  // namespace

extern "C" int mile_write_npz_batch(const char* const* paths, int32_t n_files, const char* const* member_names,
                                    const uint8_t* const* member_headers, const int32_t* header_lens,
                                    const int64_t* member_floats, int32_t n_members, const float* data, int32_t n_threads) {
  if (!paths || !member_names || !member_headers || !header_lens || !member_floats || !data || n_files < 0 || n_members <= 0)
    return -1;
  std::vector<Member> members((size_t)n_members);
  long per_file = 0;
  for (int m = 0; m < n_members; ++m) {
    members[m] = Member{std::string(member_names[m]), member_headers[m], header_lens[m], (long)member_floats[m]};
    per_file += (long)member_floats[m];
  }
  if (n_threads < 1) n_threads = 1;
  if (n_threads > n_files) n_threads = n_files > 0 ? n_files : 1;
  std::atomic<int> next(0), failed(0);
  auto work = [&]() {
    std::vector<unsigned char> out, raw, comp;
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_files) break;
      if (!build_npz(members, data + (size_t)i * per_file, out, raw, comp)) { failed.store(1); continue; }
      FILE* f = fopen(paths[i], "wb");
      if (!f) { failed.store(1); continue; }
      if (fwrite(out.data(), 1, out.size(), f) != out.size()) failed.store(1);
      if (fclose(f) != 0) failed.store(1);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  return failed.load() ? -2 : 0;
}
