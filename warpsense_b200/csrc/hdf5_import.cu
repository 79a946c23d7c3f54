// hdf5_import.cu -- host-only reader of the reference's global-map file (the subset hdf5_export.cu writes and
// libhdf5's default settings produce for such a file: version-0 superblock, version-1 object headers, groups as
// symbol tables, contiguous little-endian datasets, version-1 attributes).
//
// The reference never reads its map back (HDF5GlobalMap truncates the file on open, src/map/hdf5_global_map.cpp:5,24
// under /root/reference) -- "no resume".  ws_import_hdf5 closes that gap for this implementation: the chunks
// /map/<cx>_<cy>_<cz> go into the handle's chunk store one at a time (bounded memory), the /map attributes of
// write_meta (:208-221) and the /poses/<n>/pose rows (:175-200) are handed back.  A following ws_shift /
// ws_map_reload brings them into the device-resident local map.
#include <fcntl.h>
#include <unistd.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/warpsense_b200.h"
#include "ws_internal.h"

namespace {

struct Reader
{
  int fd = -1;
  uint64_t size = 0;
  unsigned leaf_k = 4, int_k = 16;
  ~Reader() { if (fd >= 0) ::close(fd); }
  void read(uint64_t off, void *dst, size_t n) const
  {
    if (off + n > size) throw std::runtime_error("hdf5: read past the end of the file");
    size_t done = 0;
    while (done < n)
    {
      const ssize_t r = ::pread(fd, static_cast<char *>(dst) + done, n - done, (off_t)(off + done));
      if (r <= 0) throw std::runtime_error("hdf5: read error");
      done += (size_t)r;
    }
  }
  template <typename T> T get(uint64_t off) const { T v; read(off, &v, sizeof(T)); return v; }
};

struct SymEntry { uint64_t name_off, header, btree, heap; uint32_t cache; std::string name; };
struct Link { std::string name; uint64_t header; };

SymEntry symbol_entry(const Reader &r, uint64_t off)
{
  SymEntry e;
  e.name_off = r.get<uint64_t>(off); e.header = r.get<uint64_t>(off + 8); e.cache = r.get<uint32_t>(off + 16);
  e.btree = r.get<uint64_t>(off + 24); e.heap = r.get<uint64_t>(off + 32);
  return e;
}

std::string heap_name(const Reader &r, uint64_t heap, uint64_t off)
{
  char sig[4];
  r.read(heap, sig, 4);
  if (std::memcmp(sig, "HEAP", 4) != 0) throw std::runtime_error("hdf5: bad local heap");
  const uint64_t size = r.get<uint64_t>(heap + 8), data = r.get<uint64_t>(heap + 24);
  if (off >= size) throw std::runtime_error("hdf5: heap offset");
  std::string s;
  for (uint64_t p = data + off; p < data + size; p++)
  {
    const char c = r.get<char>(p);
    if (!c) break;
    s.push_back(c);
  }
  return s;
}

void btree_links(const Reader &r, uint64_t addr, uint64_t heap, std::vector<Link> &out, int depth = 0)
{
  if (depth > 16) throw std::runtime_error("hdf5: B-tree too deep");
  char sig[4];
  r.read(addr, sig, 4);
  if (std::memcmp(sig, "TREE", 4) != 0 || r.get<uint8_t>(addr + 4) != 0) throw std::runtime_error("hdf5: bad B-tree node");
  const unsigned level = r.get<uint8_t>(addr + 5), used = r.get<uint16_t>(addr + 6);
  if (used > 2 * r.int_k) throw std::runtime_error("hdf5: B-tree node overfull");
  for (unsigned i = 0; i < used; i++)
  {
    const uint64_t child = r.get<uint64_t>(addr + 24 + 8 + 16ull * i);
    if (level > 0) { btree_links(r, child, heap, out, depth + 1); continue; }
    r.read(child, sig, 4);
    if (std::memcmp(sig, "SNOD", 4) != 0) throw std::runtime_error("hdf5: bad symbol table node");
    const unsigned n = r.get<uint16_t>(child + 6);
    if (n > 2 * r.leaf_k) throw std::runtime_error("hdf5: symbol table node overfull");
    for (unsigned k = 0; k < n; k++)
    {
      const SymEntry e = symbol_entry(r, child + 8 + 40ull * k);
      Link l; l.name = heap_name(r, heap, e.name_off); l.header = e.header;
      out.push_back(l);
    }
  }
}

struct Object
{
  bool is_group = false;
  uint64_t btree = 0, heap = 0;
  // dataset
  uint64_t data_addr = 0, data_bytes = 0, n_elem = 0;
  int type_class = -1, type_size = 0;
  struct Attr { std::string name; int type_class; uint32_t raw; };
  std::vector<Attr> attrs;
};

Object read_object(const Reader &r, uint64_t addr)
{
  Object o;
  if (r.get<uint8_t>(addr) != 1) throw std::runtime_error("hdf5: object header version");
  const unsigned nmsg = r.get<uint16_t>(addr + 2);
  const uint32_t size = r.get<uint32_t>(addr + 8);
  uint64_t p = addr + 16;
  const uint64_t end = p + size;
  std::vector<uint8_t> d;
  for (unsigned m = 0; m < nmsg && p < end; m++)
  {
    const unsigned type = r.get<uint16_t>(p), msize = r.get<uint16_t>(p + 2);
    d.resize(msize);
    if (msize) r.read(p + 8, d.data(), msize);
    auto u64at = [&](size_t off) { uint64_t v; std::memcpy(&v, d.data() + off, 8); return v; };
    if (type == 0x0011 && msize >= 16) { o.is_group = true; o.btree = u64at(0); o.heap = u64at(8); }
    else if (type == 0x0001 && msize >= 8)
    {
      const unsigned rank = d[1];
      o.n_elem = 1;
      for (unsigned k = 0; k < rank; k++) o.n_elem *= u64at(8 + 8 * k);
    }
    else if (type == 0x0003 && msize >= 8) { o.type_class = d[0] & 0x0F; std::memcpy(&o.type_size, d.data() + 4, 4); }
    else if (type == 0x0008 && msize >= 18)
    {
      if (d[0] != 3 || d[1] != 1) throw std::runtime_error("hdf5: only contiguous version-3 layouts are supported");
      o.data_addr = u64at(2); o.data_bytes = u64at(10);
    }
    else if (type == 0x000C && msize >= 8)
    {
      uint16_t nsz, tsz, ssz;
      std::memcpy(&nsz, d.data() + 2, 2); std::memcpy(&tsz, d.data() + 4, 2); std::memcpy(&ssz, d.data() + 6, 2);
      auto pad = [](size_t n) { return (n + 7) & ~(size_t)7; };
      size_t q = 8;
      Object::Attr a;
      a.name.assign(reinterpret_cast<const char *>(d.data() + q), strnlen(reinterpret_cast<const char *>(d.data() + q), nsz));
      q += pad(nsz);
      a.type_class = d[q] & 0x0F;
      q += pad(tsz) + pad(ssz);
      a.raw = 0;
      if (q + 4 <= d.size()) std::memcpy(&a.raw, d.data() + q, 4);
      o.attrs.push_back(a);
    }
    p += 8 + msize;
  }
  return o;
}

}  // namespace

extern "C" int ws_import_hdf5(ws_handle *h, const char *path, ws_map_meta *meta, float *poses7, int64_t poses_cap,
                              int64_t *n_chunks, int64_t *n_poses)
{
  if (!h || !path) return WS_ERR_INVALID;
  try
  {
    Reader r;
    r.fd = ::open(path, O_RDONLY);
    if (r.fd < 0) throw std::runtime_error(std::string("cannot open ") + path);
    r.size = (uint64_t)::lseek(r.fd, 0, SEEK_END);
    uint8_t sb[24];
    r.read(0, sb, 24);
    const uint8_t sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
    if (std::memcmp(sb, sig, 8) != 0) throw std::runtime_error("hdf5: bad signature");
    if (sb[8] != 0 || sb[13] != 8 || sb[14] != 8) throw std::runtime_error("hdf5: unsupported superblock");
    std::memcpy(&r.leaf_k, sb + 16, 2); r.leaf_k &= 0xFFFF;
    std::memcpy(&r.int_k, sb + 18, 2); r.int_k &= 0xFFFF;
    const SymEntry root = symbol_entry(r, 56);
    const Object root_o = read_object(r, root.header);
    if (!root_o.is_group) throw std::runtime_error("hdf5: the root is not a group");
    std::vector<Link> top;
    btree_links(r, root_o.btree, root_o.heap, top);
    int64_t chunks = 0, poses = 0;
    std::vector<uint32_t> buf((size_t)64 * 64 * 64);
    for (const Link &t : top)
    {
      const Object g = read_object(r, t.header);
      if (!g.is_group) continue;
      std::vector<Link> links;
      btree_links(r, g.btree, g.heap, links);
      if (t.name == "map")
      {
        if (meta)
          for (const auto &a : g.attrs)
          {
            int32_t iv; float fv;
            std::memcpy(&iv, &a.raw, 4); std::memcpy(&fv, &a.raw, 4);
            if (a.name == "tau") meta->tau = iv;
            else if (a.name == "map_size_x") meta->map_size[0] = iv;
            else if (a.name == "map_size_y") meta->map_size[1] = iv;
            else if (a.name == "map_size_z") meta->map_size[2] = iv;
            else if (a.name == "max_distance") meta->max_distance = fv;
            else if (a.name == "map_resolution") meta->map_resolution = iv;
            else if (a.name == "max_weight") meta->max_weight = iv;
          }
        for (const Link &l : links)
        {
          int cx, cy, cz;
          if (std::sscanf(l.name.c_str(), "%d_%d_%d", &cx, &cy, &cz) != 3) continue;       // tag_from_chunk_pos, :46-51
          const Object d = read_object(r, l.header);
          if (d.is_group || d.n_elem != buf.size() || d.type_size != 4 || d.data_bytes != buf.size() * 4)
            throw std::runtime_error("hdf5: /map/" + l.name + " is not a 64^3 uint32 dataset");
          r.read(d.data_addr, buf.data(), buf.size() * 4);
          std::memcpy(const_cast<uint32_t *>(ws_store_chunk(h, cx, cy, cz)), buf.data(), buf.size() * 4);
          chunks++;
        }
      }
      else if (t.name == "poses")
      {
        // /poses/<n>/pose, n ascending
        std::vector<std::pair<long, uint64_t>> order;
        for (const Link &l : links) order.push_back({ std::strtol(l.name.c_str(), nullptr, 10), l.header });
        std::sort(order.begin(), order.end());
        for (const auto &po : order)
        {
          const Object pg = read_object(r, po.second);
          if (!pg.is_group) continue;
          std::vector<Link> pl;
          btree_links(r, pg.btree, pg.heap, pl);
          for (const Link &l : pl)
          {
            if (l.name != "pose") continue;
            const Object d = read_object(r, l.header);
            if (d.n_elem != 7 || d.type_size != 4) throw std::runtime_error("hdf5: a pose is not 7 float32");
            if (poses7 && poses < poses_cap) r.read(d.data_addr, poses7 + 7 * poses, 28);
            poses++;
          }
        }
      }
    }
    if (n_chunks) *n_chunks = chunks;
    if (n_poses) *n_poses = poses;
    return WS_OK;
  }
  catch (const std::exception &e)
  {
    h->last_error = e.what();
    return WS_ERR_STATE;
  }
}
