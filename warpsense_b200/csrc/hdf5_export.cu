// hdf5_export.cu -- host-only, dependency-free writer of the reference's global-map file.
//
// Replaces HDF5GlobalMap's file side (src/map/hdf5_global_map.cpp under /root/reference, which goes through
// HighFive/libhdf5 -- neither exists in this image): the same logical layout
//     /map                      group, attributes tau, map_size_{x,y,z}, max_distance, map_resolution, max_weight
//                               (write_meta, :208-221)
//     /map/<cx>_<cy>_<cz>       1-D dataset of 64^3 uint32 raw TSDF entries, index x*4096 + y*64 + z
//                               (tag_from_chunk_pos :46-51, index_from_pos :53-57, createDataSet :120,:170)
//     /poses/<n>/pose           1-D dataset of 7 float32: x y z (scaled, rounded to 1e-3) qx qy qz qw
//                               (write_pose :175-200)
// written directly in the HDF5 file format as libhdf5's default (earliest) settings would: version-0
// superblock, version-1 object headers, groups as symbol tables (v1 B-tree of SNOD leaves + local heap),
// contiguous little-endian datasets, version-1 attribute messages.  One pass: the 1 MiB chunk payloads are
// streamed to disk as they come, only headers / B-trees / heaps are assembled in memory.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/warpsense_b200.h"
#include "ws_internal.h"

namespace {

const uint64_t UNDEF = ~0ull;
const int LEAF_K = 4;        // symbol table node holds 2K entries      (libhdf5 defaults)
const int INTERNAL_K = 16;   // B-tree node holds 2K children

// Append-only byte sink.  Small pieces (headers, B-trees, heaps) collect in memory; with a file attached
// (streaming mode) the big dataset payloads go straight to disk and the memory part is flushed in front of
// them, so the file never has to exist in memory as a whole.
struct Buf
{
  std::vector<uint8_t> b;
  FILE *fp = nullptr;          // streaming: everything before `base` is already on disk
  uint64_t base = 0;
  uint64_t size() const { return base + b.size(); }
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { for (int i = 0; i < 2; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void u32(uint32_t v) { for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; i++) b.push_back((uint8_t)(v >> (8 * i))); }
  void bytes(const void *p, size_t n) { const uint8_t *q = static_cast<const uint8_t *>(p); b.insert(b.end(), q, q + n); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  void align8() { while (size() & 7) b.push_back(0); }
  void flush()
  {
    if (!fp || b.empty()) return;
    if (std::fseek(fp, (long)base, SEEK_SET) != 0 || std::fwrite(b.data(), 1, b.size(), fp) != b.size())
      throw std::runtime_error("short write");
    base += b.size();
    b.clear();
  }
  // a large payload: in memory without a file, else written through
  void payload(const void *p, size_t n)
  {
    if (!fp) { bytes(p, n); return; }
    flush();
    if (std::fseek(fp, (long)base, SEEK_SET) != 0 || std::fwrite(p, 1, n, fp) != n) throw std::runtime_error("short write");
    base += n;
  }
  void patch(uint64_t off, const void *p, size_t n)
  {
    if (off >= base) { std::memcpy(b.data() + (off - base), p, n); return; }
    if (off + n > base) throw std::runtime_error("patch straddles the flushed part");
    if (std::fseek(fp, (long)off, SEEK_SET) != 0 || std::fwrite(p, 1, n, fp) != n) throw std::runtime_error("short write");
  }
  void put_u64_at(uint64_t off, uint64_t v)
  {
    uint8_t t[8];
    for (int i = 0; i < 8; i++) t[i] = (uint8_t)(v >> (8 * i));
    patch(off, t, 8);
  }
};

// ---- header messages ------------------------------------------------------------------------------
struct Msg
{
  uint16_t type;
  uint8_t flags;               // bit 0: constant (libhdf5 marks a dataset's dataspace/datatype/fill value so)
  std::vector<uint8_t> data;   // padded to a multiple of 8
};

Msg make_msg(uint16_t type, Buf &d, uint8_t flags = 0)
{
  d.align8();
  Msg m; m.type = type; m.flags = flags; m.data = d.b;
  return m;
}

// datatype message bodies (class + version 1)
void dt_fixed32(Buf &d, bool is_signed)
{
  d.u8(0x10); d.u8(is_signed ? 0x08 : 0x00); d.u8(0); d.u8(0);   // class 0 (fixed point), little endian
  d.u32(4);                                                      // size
  d.u16(0); d.u16(32);                                           // bit offset, precision
}
void dt_float32(Buf &d)
{
  d.u8(0x11); d.u8(0x20); d.u8(0x1f); d.u8(0x00);                // class 1, LE, implied mantissa msb, sign bit 31
  d.u32(4);
  d.u16(0); d.u16(32);                                           // bit offset, precision
  d.u8(23); d.u8(8); d.u8(0); d.u8(23);                          // exponent location/size, mantissa location/size
  d.u32(127);                                                    // exponent bias
}

Msg msg_dataspace_1d(uint64_t n)
{
  Buf d; d.u8(1); d.u8(1); d.u8(0); d.u8(0); d.u32(0); d.u64(n);
  return make_msg(0x0001, d, 1);
}
Msg msg_datatype(int kind)   // 0 uint32, 1 int32, 2 float32
{
  Buf d;
  if (kind == 2) dt_float32(d); else dt_fixed32(d, kind == 1);
  return make_msg(0x0003, d, 1);
}
Msg msg_fill_value()
{
  Buf d; d.u8(2); d.u8(2); d.u8(2); d.u8(1); d.u32(0);           // v2, alloc late, write if set, defined, size 0
  return make_msg(0x0005, d, 1);
}
Msg msg_layout_contiguous(uint64_t addr, uint64_t bytes)
{
  Buf d; d.u8(3); d.u8(1); d.u64(addr); d.u64(bytes);
  return make_msg(0x0008, d);
}
Msg msg_symbol_table(uint64_t btree, uint64_t heap)
{
  Buf d; d.u64(btree); d.u64(heap);
  return make_msg(0x0011, d);
}
// version-1 attribute, scalar dataspace
Msg msg_attribute(const std::string &name, int kind, const void *value)
{
  Buf dt; if (kind == 2) dt_float32(dt); else dt_fixed32(dt, kind == 1);
  const uint16_t dt_size = (uint16_t)dt.size();
  Buf ds; ds.u8(1); ds.u8(0); ds.u8(0); ds.u8(0); ds.u32(0);      // rank 0
  const uint16_t ds_size = (uint16_t)ds.size();
  Buf d;
  d.u8(1); d.u8(0);
  d.u16((uint16_t)(name.size() + 1)); d.u16(dt_size); d.u16(ds_size);
  d.bytes(name.c_str(), name.size() + 1); d.align8();
  d.bytes(dt.b.data(), dt.size()); d.align8();
  d.bytes(ds.b.data(), ds.size()); d.align8();
  d.bytes(value, 4);
  return make_msg(0x000C, d);
}

// version-1 object header; returns its address
uint64_t write_object_header(Buf &f, const std::vector<Msg> &msgs)
{
  f.align8();
  const uint64_t addr = f.size();
  uint32_t body = 0;
  for (const Msg &m : msgs) body += 8 + (uint32_t)m.data.size();
  f.u8(1); f.u8(0); f.u16((uint16_t)msgs.size()); f.u32(1); f.u32(body);
  f.u32(0);                                                      // pad the 12-byte prefix to 16
  for (const Msg &m : msgs)
  {
    f.u16(m.type); f.u16((uint16_t)m.data.size()); f.u8(m.flags); f.u8(0); f.u8(0); f.u8(0);
    f.bytes(m.data.data(), m.data.size());
  }
  return addr;
}

struct Entry            // one link of a group
{
  std::string name;
  uint64_t header = 0;  // object header address
  bool is_group = false;
  uint64_t btree = 0, heap = 0;   // cached in the symbol table entry of a group (cache type 1)
};

void write_symbol_entry(Buf &f, uint64_t name_off, const Entry &e)
{
  f.u64(name_off); f.u64(e.header);
  f.u32(e.is_group ? 1u : 0u); f.u32(0);
  if (e.is_group) { f.u64(e.btree); f.u64(e.heap); } else { f.u64(0); f.u64(0); }
}

// A group: local heap (names), symbol table nodes, B-tree, object header with the symbol table message.
// `entries` must be sorted by name (strcmp).  extra: further messages for the group's header (attributes).
Entry write_group(Buf &f, const std::string &name, std::vector<Entry> entries, const std::vector<Msg> &extra)
{
  std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return std::strcmp(a.name.c_str(), b.name.c_str()) < 0; });
  // ---- local heap data: "" at offset 0, then every name, 8-aligned; a free block closes the segment
  Buf hd;
  hd.zeros(8);
  std::vector<uint64_t> name_off(entries.size());
  for (size_t i = 0; i < entries.size(); i++)
  {
    name_off[i] = hd.size();
    hd.bytes(entries[i].name.c_str(), entries[i].name.size() + 1);
    hd.align8();
  }
  const uint64_t free_off = hd.size();
  hd.u64(1);                    // next free block: none (H5HL_FREE_NULL)
  hd.u64(32);                   // size of this free block
  hd.zeros(16);
  f.align8();
  const uint64_t heap_addr = f.size();
  f.bytes("HEAP", 4); f.u8(0); f.u8(0); f.u8(0); f.u8(0);
  f.u64(hd.size()); f.u64(free_off); f.u64(heap_addr + 32);
  f.bytes(hd.b.data(), hd.size());

  // ---- symbol table nodes (leaves), 2*LEAF_K entries each, always allocated in full
  struct Leaf { uint64_t addr; uint64_t last_name_off; };
  std::vector<Leaf> level;
  const size_t per_leaf = 2 * LEAF_K;
  size_t pos = 0;
  if (entries.empty())
  {
    // an empty group: a B-tree root without children, no symbol table node
    f.align8();
    const uint64_t bt = f.size();
    f.bytes("TREE", 4); f.u8(0); f.u8(0); f.u16(0); f.u64(UNDEF); f.u64(UNDEF);
    f.zeros(8 + 2 * INTERNAL_K * 16);
    std::vector<Msg> msgs0;
    msgs0.push_back(msg_symbol_table(bt, heap_addr));
    for (const Msg &m : extra) msgs0.push_back(m);
    Entry g0;
    g0.name = name; g0.is_group = true; g0.btree = bt; g0.heap = heap_addr;
    g0.header = write_object_header(f, msgs0);
    return g0;
  }
  do
  {
    const size_t cnt = std::min(per_leaf, entries.size() - pos);
    f.align8();
    Leaf lf; lf.addr = f.size(); lf.last_name_off = cnt ? name_off[pos + cnt - 1] : 0;
    f.bytes("SNOD", 4); f.u8(1); f.u8(0); f.u16((uint16_t)cnt);
    for (size_t i = 0; i < per_leaf; i++)
    {
      if (i < cnt) write_symbol_entry(f, name_off[pos + i], entries[pos + i]);
      else f.zeros(40);
    }
    level.push_back(lf);
    pos += cnt;
  } while (pos < entries.size());

  // ---- B-tree: level 0 nodes point at the leaves; more levels while a level has more than one node
  int depth = 0;
  for (;;)
  {
    std::vector<Leaf> next;
    const size_t per_node = 2 * INTERNAL_K;
    for (size_t p = 0; p < level.size(); p += per_node)
    {
      const size_t cnt = std::min(per_node, level.size() - p);
      f.align8();
      Leaf nd; nd.addr = f.size(); nd.last_name_off = level[p + cnt - 1].last_name_off;
      f.bytes("TREE", 4); f.u8(0); f.u8((uint8_t)depth); f.u16((uint16_t)cnt);
      f.u64(UNDEF); f.u64(UNDEF);                         // siblings are patched below
      // key 0: the name before everything in this node ("" for the leftmost node)
      f.u64(p == 0 ? 0 : level[p - 1].last_name_off);
      for (size_t i = 0; i < per_node; i++)
      {
        if (i < cnt) { f.u64(level[p + i].addr); f.u64(level[p + i].last_name_off); }
        else { f.u64(0); f.u64(0); }
      }
      next.push_back(nd);
    }
    for (size_t i = 0; i < next.size(); i++)              // sibling links of this level
    {
      if (i > 0) f.put_u64_at(next[i].addr + 8, next[i - 1].addr);
      if (i + 1 < next.size()) f.put_u64_at(next[i].addr + 16, next[i + 1].addr);
    }
    level = next;
    depth++;
    if (level.size() == 1) break;
  }
  const uint64_t btree_addr = level[0].addr;

  std::vector<Msg> msgs;
  msgs.push_back(msg_symbol_table(btree_addr, heap_addr));
  for (const Msg &m : extra) msgs.push_back(m);
  Entry g;
  g.name = name; g.is_group = true; g.btree = btree_addr; g.heap = heap_addr;
  g.header = write_object_header(f, msgs);
  return g;
}

Entry write_dataset(Buf &f, const std::string &name, int kind, const void *data, uint64_t n)
{
  f.align8();
  const uint64_t data_addr = f.size();
  f.payload(data, n * 4);
  std::vector<Msg> msgs;
  msgs.push_back(msg_dataspace_1d(n));
  msgs.push_back(msg_datatype(kind));
  msgs.push_back(msg_fill_value());
  msgs.push_back(msg_layout_contiguous(data_addr, n * 4));
  Entry e; e.name = name; e.header = write_object_header(f, msgs);
  return e;
}

}  // namespace

namespace {

struct ChunkRef { int x, y, z; const uint32_t *data; };

// `fetch(i)` returns the 64^3 entries of chunk i (valid until the next call): the chunks are streamed to disk
// one at a time, only headers / B-trees / heaps are assembled in memory.
template <typename Fetch>
void write_map_file(const char *path, const ws_map_meta *meta, const std::vector<ChunkRef> &chunk_refs, Fetch &&fetch,
                    const float *poses7, int64_t n_poses)
{
  FILE *fp = std::fopen(path, "w+b");
  if (!fp) throw std::runtime_error(std::string("cannot open ") + path);
  struct Closer { FILE *fp; ~Closer() { if (fp) std::fclose(fp); } } closer{ fp };
  Buf f;
  f.fp = fp;
  f.zeros(96);                                            // superblock, filled in at the end
  // /map/<cx>_<cy>_<cz>
  std::vector<Entry> chunks;
  for (size_t i = 0; i < chunk_refs.size(); i++)
  {
    const ChunkRef &c = chunk_refs[i];
    const std::string tag = std::to_string(c.x) + "_" + std::to_string(c.y) + "_" + std::to_string(c.z);   // :46-51
    chunks.push_back(write_dataset(f, tag, 0, fetch(i), (uint64_t)64 * 64 * 64));
  }
  std::vector<Msg> attrs;                                 // write_meta, :208-221
  attrs.push_back(msg_attribute("tau", 1, &meta->tau));
  attrs.push_back(msg_attribute("map_size_x", 1, &meta->map_size[0]));
  attrs.push_back(msg_attribute("map_size_y", 1, &meta->map_size[1]));
  attrs.push_back(msg_attribute("map_size_z", 1, &meta->map_size[2]));
  attrs.push_back(msg_attribute("max_distance", 2, &meta->max_distance));
  attrs.push_back(msg_attribute("map_resolution", 1, &meta->map_resolution));
  attrs.push_back(msg_attribute("max_weight", 1, &meta->max_weight));
  const Entry map_g = write_group(f, "map", chunks, attrs);
  // /poses/<n>/pose
  std::vector<Entry> pose_groups;
  for (int64_t i = 0; i < n_poses; i++)
  {
    std::vector<Entry> one;
    one.push_back(write_dataset(f, "pose", 2, poses7 + 7 * i, 7));
    pose_groups.push_back(write_group(f, std::to_string(i), one, {}));
  }
  const Entry poses_g = write_group(f, "poses", pose_groups, {});
  std::vector<Entry> top;
  top.push_back(map_g); top.push_back(poses_g);
  const Entry root = write_group(f, "", top, {});
  f.align8();

  // ---- superblock, version 0
  Buf s;
  const uint8_t sig[8] = { 0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n' };
  s.bytes(sig, 8);
  s.u8(0); s.u8(0); s.u8(0); s.u8(0); s.u8(0);            // superblock, free space, root entry, reserved, shared header versions
  s.u8(8); s.u8(8); s.u8(0);                              // size of offsets, size of lengths, reserved
  s.u16(LEAF_K); s.u16(INTERNAL_K);
  s.u32(0);                                               // file consistency flags
  s.u64(0); s.u64(UNDEF); s.u64(f.size()); s.u64(UNDEF);  // base, free-space info, end of file, driver info
  write_symbol_entry(s, 0, root);                         // root group symbol table entry
  if (s.size() != 96) throw std::runtime_error("superblock size");
  f.flush();
  f.patch(0, s.b.data(), 96);
  if (std::fflush(fp) != 0) throw std::runtime_error("short write");
}

}  // namespace

extern "C" int ws_export_hdf5(ws_handle *h, const char *path, const ws_map_meta *meta, const float *poses7, int64_t n_poses)
{
  if (!h || !path || !meta || n_poses < 0 || (n_poses > 0 && !poses7)) return WS_ERR_INVALID;
  try
  {
    std::vector<ChunkRef> refs;
    std::vector<unsigned long long> keys = ws_store_keys(h);
    for (unsigned long long k : keys)
    {
      ChunkRef c;
      c.x = (int)((k >> 42) & 0x1FFFFF) - (1 << 20); c.y = (int)((k >> 21) & 0x1FFFFF) - (1 << 20);
      c.z = (int)(k & 0x1FFFFF) - (1 << 20);
      c.data = nullptr;
      refs.push_back(c);
    }
    // one chunk at a time through the bounded store (spilled chunks are read back in turn)
    write_map_file(path, meta, refs, [&](size_t i) { return ws_store_chunk(h, refs[i].x, refs[i].y, refs[i].z); }, poses7, n_poses);
    return WS_OK;
  }
  catch (const std::exception &e)
  {
    h->last_error = e.what();
    return WS_ERR_STATE;
  }
}


// the same file from caller-held chunks (no map handle, no GPU): chunk_xyz[n_chunks][3], chunk_data[n_chunks][64^3]
extern "C" int ws_hdf5_write_chunks(const char *path, const ws_map_meta *meta, const int32_t *chunk_xyz,
                                    const uint32_t *chunk_data, int64_t n_chunks, const float *poses7, int64_t n_poses)
{
  if (!path || !meta || n_chunks < 0 || n_poses < 0 || (n_chunks > 0 && (!chunk_xyz || !chunk_data)) || (n_poses > 0 && !poses7))
    return WS_ERR_INVALID;
  try
  {
    std::vector<ChunkRef> refs;
    for (int64_t i = 0; i < n_chunks; i++)
    {
      ChunkRef c;
      c.x = chunk_xyz[3 * i]; c.y = chunk_xyz[3 * i + 1]; c.z = chunk_xyz[3 * i + 2];
      c.data = chunk_data + (size_t)i * 64 * 64 * 64;
      refs.push_back(c);
    }
    write_map_file(path, meta, refs, [&](size_t i) { return refs[i].data; }, poses7, n_poses);
    return WS_OK;
  }
  catch (const std::exception &) { return WS_ERR_STATE; }
}
