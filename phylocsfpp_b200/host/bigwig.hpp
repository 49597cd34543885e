// bigwig.hpp — bigWig reader and writer of the phylocsf_b200 host.
//
// The reference reads its finished tracks through libBigWig (an un-vendored dependency: `#include <bigWig.h>`,
// src/phylocsf++annotate_with_tracks.hpp:10-14) and leaves writing them to UCSC's external wigToBigWig.  Neither is in
// this image, so both directions are restated here from the published file format (Kent et al. 2010, "BigWig and
// BigBed", supplement: common header, chromosome B+ tree, data sections, R-tree index, zoom levels).  The reader is
// pinned by reproducing the reference's expected annotate-with-tracks output from the reference's own example/tracks/*.bw
// (tests/test_host_annotate.py); the writer is pinned by round trips through that reader.
//
//   BigWigReader::values(chrom, begin, end, out)  ==  libBigWig's bwGetValues(fp, chrom, begin, end, includeNA = 1):
//     one float per base of [begin, end), NaN where the file has no value; false (libBigWig: NULL) when the chromosome is
//     unknown or the range is empty after clamping `end` to the chromosome length — the length of `out` stays end - begin.
//   BigWigWriter: fixedStep / varStep runs in, sections of <= 1024 items (zlib), chromosome tree, R-tree, total summary and
//     zoom levels out — the layout wigToBigWig produces, so UCSC tools and libBigWig read the files.
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace host {

constexpr uint32_t BW_MAGIC = 0x888FFC26u, BW_CHROM_TREE_MAGIC = 0x78CA8C91u, BW_RTREE_MAGIC = 0x2468ACE0u;

// --------------------------------------------------------------------------------------------------------------- reader
class BigWigReader {
public:
    struct Chrom { std::string name; uint32_t id, len; };

    BigWigReader() = default;
    BigWigReader(const BigWigReader &) = delete;
    BigWigReader &operator=(const BigWigReader &) = delete;
    ~BigWigReader() { close(); }

    bool open(const std::string &path, std::string &err) {
        close();
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) { err = "cannot open " + path; return false; }
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size < 64) { ::close(fd); err = path + " is not a bigWig file (too short)"; return false; }
        size_ = (size_t)st.st_size;
        void *m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (m == MAP_FAILED) { err = "cannot map " + path; return false; }
        mem_ = static_cast<const uint8_t *>(m);
        if (u32(0) != BW_MAGIC) { err = path + " is not a little-endian bigWig file (wrong magic)"; close(); return false; }
        version_ = u16(4);
        n_zoom_ = u16(6);
        const uint64_t chrom_tree = u64(8);
        data_off_ = u64(16);
        index_off_ = u64(24);
        summary_off_ = u64(44);
        compressed_ = u32(52) > 0;
        buf_.resize(std::max<uint32_t>(u32(52), 1u << 16));
        if (chrom_tree + 32 > size_ || index_off_ + 48 > size_ || u32(chrom_tree) != BW_CHROM_TREE_MAGIC || u32(index_off_) != BW_RTREE_MAGIC) {
            err = path + ": damaged bigWig header";
            close();
            return false;
        }
        key_size_ = u32(chrom_tree + 8);
        if (!walk_chrom_tree(chrom_tree + 32, 0)) { err = path + ": damaged chromosome tree"; close(); return false; }
        for (const Chrom &c : chroms_) by_name_[c.name] = c;
        return true;
    }

    void close() {
        if (mem_) munmap(const_cast<uint8_t *>(mem_), size_);
        mem_ = nullptr;
        chroms_.clear();
        by_name_.clear();
    }

    const std::vector<Chrom> &chroms() const { return chroms_; }          // bwReadChromList order (tree order = sorted by name)
    const Chrom *find(const std::string &name) const { auto it = by_name_.find(name); return it == by_name_.end() ? nullptr : &it->second; }
    int zoom_levels() const { return n_zoom_; }
    int version() const { return version_; }
    uint64_t n_sections() const { return u64(data_off_); }

    struct Summary { uint64_t bases_covered; double min, max, sum, sum_squares; };
    Summary summary() const {
        Summary s{};
        if (summary_off_ && summary_off_ + 40 <= size_) { s.bases_covered = u64(summary_off_); s.min = f64(summary_off_ + 8); s.max = f64(summary_off_ + 16); s.sum = f64(summary_off_ + 24); s.sum_squares = f64(summary_off_ + 32); }
        return s;
    }

    // bwGetValues(..., includeNA = 1).  `out` gets end - begin floats.
    bool values(const std::string &chrom, uint32_t begin, uint32_t end, std::vector<float> &out) {
        out.clear();
        const Chrom *c = find(chrom);
        if (!c) return false;
        const uint32_t cend = std::min(end, c->len);
        if (begin >= cend) return false;
        out.assign((size_t)end - begin, std::numeric_limits<float>::quiet_NaN());
        bool ok = true;
        walk_rtree(index_off_ + 48, c->id, begin, cend, [&](uint64_t off, uint64_t sz) { ok = ok && section_values(off, sz, c->id, begin, end, out.data()); });
        return ok;
    }

    // Every interval of the file in file order: f(chrom_id, start, end, value) — what bigWigToBedGraph prints.
    template <class F>
    bool for_each_interval(F &&f) {
        bool ok = true;
        walk_rtree(index_off_ + 48, 0, 0, 0, [&](uint64_t off, uint64_t sz) { ok = ok && section_intervals(off, sz, f); }, true);
        return ok;
    }

private:
    const uint8_t *mem_ = nullptr;
    size_t size_ = 0;
    uint16_t version_ = 0, n_zoom_ = 0;
    uint64_t data_off_ = 0, index_off_ = 0, summary_off_ = 0;
    uint32_t key_size_ = 0;
    bool compressed_ = false;
    std::vector<uint8_t> buf_;
    std::vector<Chrom> chroms_;
    std::map<std::string, Chrom> by_name_;

    template <class T> T rd(uint64_t o) const { T v; memcpy(&v, mem_ + o, sizeof(T)); return v; }
    uint16_t u16(uint64_t o) const { return rd<uint16_t>(o); }
    uint32_t u32(uint64_t o) const { return rd<uint32_t>(o); }
    uint64_t u64(uint64_t o) const { return rd<uint64_t>(o); }
    double f64(uint64_t o) const { return rd<double>(o); }

    bool walk_chrom_tree(uint64_t node, int depth) {
        if (node + 4 > size_ || depth > 32) return false;
        const bool leaf = mem_[node] != 0;
        const uint16_t n = u16(node + 2);
        const uint64_t item = key_size_ + 8;
        if (node + 4 + n * item > size_) return false;
        for (uint16_t i = 0; i < n; ++i) {
            const uint64_t o = node + 4 + i * item;
            if (leaf) {
                const char *k = reinterpret_cast<const char *>(mem_ + o);
                chroms_.push_back(Chrom{std::string(k, strnlen(k, key_size_)), u32(o + key_size_), u32(o + key_size_ + 4)});
            } else if (!walk_chrom_tree(u64(o + key_size_), depth + 1)) {
                return false;
            }
        }
        return true;
    }

    // (chrom, base) pairs compare lexicographically; a node overlaps [ (id,begin), (id,end) ) like this:
    static bool overlaps(uint32_t id, uint32_t begin, uint32_t end, uint32_t c0, uint32_t b0, uint32_t c1, uint32_t b1) {
        const bool starts_before_end = c0 < id || (c0 == id && b0 < end);
        const bool ends_after_begin = c1 > id || (c1 == id && b1 > begin);
        return starts_before_end && ends_after_begin;
    }

    template <class F>
    void walk_rtree(uint64_t node, uint32_t id, uint32_t begin, uint32_t end, F &&f, bool all = false, int depth = 0) {
        if (node + 4 > size_ || depth > 32) return;
        const bool leaf = mem_[node] != 0;
        const uint16_t n = u16(node + 2);
        const uint64_t item = leaf ? 32 : 24;
        if (node + 4 + n * item > size_) return;
        for (uint16_t i = 0; i < n; ++i) {
            const uint64_t o = node + 4 + i * item;
            if (!all && !overlaps(id, begin, end, u32(o), u32(o + 4), u32(o + 8), u32(o + 12))) continue;
            if (leaf) f(u64(o + 16), u64(o + 24));
            else walk_rtree(u64(o + 16), id, begin, end, f, all, depth + 1);
        }
    }

    // A data section, decompressed if the file is compressed: pointer + size.
    bool section(uint64_t off, uint64_t sz, const uint8_t *&p, size_t &n) {
        if (off + sz > size_) return false;
        if (!compressed_) { p = mem_ + off; n = sz; return n >= 24; }
        for (;;) {
            uLongf dn = buf_.size();
            const int rc = uncompress(buf_.data(), &dn, mem_ + off, sz);
            if (rc == Z_OK) { p = buf_.data(); n = dn; return n >= 24; }
            if (rc != Z_BUF_ERROR || buf_.size() > (1u << 28)) return false;
            buf_.resize(buf_.size() * 2);
        }
    }

    template <class F>
    bool section_items(const uint8_t *p, size_t n, F &&f) {
        uint32_t h[5];
        memcpy(h, p, 20);
        const uint32_t chrom = h[0], step = h[3], span = h[4];
        uint32_t start = h[1];
        const uint8_t type = p[20];
        uint16_t count;
        memcpy(&count, p + 22, 2);
        const size_t item = type == 1 ? 12 : type == 2 ? 8 : 4;
        if (type < 1 || type > 3 || 24 + count * item > n) return false;
        const uint8_t *q = p + 24;
        for (uint16_t i = 0; i < count; ++i, q += item) {
            uint32_t s, e;
            float v;
            if (type == 1) { memcpy(&s, q, 4); memcpy(&e, q + 4, 4); memcpy(&v, q + 8, 4); }
            else if (type == 2) { memcpy(&s, q, 4); memcpy(&v, q + 4, 4); e = s + span; }
            else { memcpy(&v, q, 4); s = start; e = s + span; start += step; }
            f(chrom, s, e, v);
        }
        return true;
    }

    bool section_values(uint64_t off, uint64_t sz, uint32_t id, uint32_t begin, uint32_t end, float *out) {
        const uint8_t *p;
        size_t n;
        if (!section(off, sz, p, n)) return false;
        return section_items(p, n, [&](uint32_t chrom, uint32_t s, uint32_t e, float v) {
            if (chrom != id) return;
            for (uint32_t j = std::max(s, begin); j < std::min(e, end); ++j) out[j - begin] = v;
        });
    }

    template <class F>
    bool section_intervals(uint64_t off, uint64_t sz, F &&f) {
        const uint8_t *p;
        size_t n;
        if (!section(off, sz, p, n)) return false;
        return section_items(p, n, f);
    }
};

// --------------------------------------------------------------------------------------------------------------- writer
// Usage: BigWigWriter w; w.open(path, chroms /* name -> length */); then, per run of a wig file in file order (chromosomes
// grouped, positions ascending): w.add_fixed_step(chrom, start0 /* 0-based */, step, span, values, n); finally w.finish().
class BigWigWriter {
public:
    static constexpr uint32_t ITEMS_PER_SLOT = 1024, BLOCK_SIZE = 256;

    bool open(const std::string &path, const std::vector<std::pair<std::string, uint32_t>> &chroms, std::string &err, bool compress = true) {
        fo_ = fopen(path.c_str(), "wb");
        if (!fo_) { err = "cannot create " + path; return false; }
        compress_ = compress;
        chroms_ = chroms;
        std::sort(chroms_.begin(), chroms_.end());                  // the B+ tree is ordered by name; ids follow that order
        for (size_t i = 0; i < chroms_.size(); ++i) id_[chroms_[i].first] = (uint32_t)i;
        key_size_ = 1;
        for (auto &c : chroms_) key_size_ = std::max<uint32_t>(key_size_, (uint32_t)c.first.size());
        // fixed-size head: header, room for up to MAX_ZOOM zoom headers, total summary, chromosome tree; then the data
        std::vector<uint8_t> head(64 + 24 * MAX_ZOOM, 0);
        put(head);
        summary_off_ = pos_;
        put(std::vector<uint8_t>(40, 0));
        chrom_tree_off_ = pos_;
        write_chrom_tree();
        data_off_ = pos_;
        put64(0);                                                   // section count, patched in finish()
        return true;
    }

    // One wig run.  start0 is 0-based (a wig "start=" minus one).
    bool add_fixed_step(const std::string &chrom, uint32_t start0, uint32_t step, uint32_t span, const float *v, size_t n, std::string &err) {
        auto it = id_.find(chrom);
        if (it == id_.end()) { err = "chromosome " + chrom + " is not in the chromosome list"; return false; }
        const uint32_t id = it->second, len = chroms_[id].second;
        if (n && (uint64_t)start0 + (uint64_t)(n - 1) * step + span > len) { err = "run on " + chrom + " ends behind the chromosome"; return false; }
        if (!sections_.empty() && n && (id < last_id_ || (id == last_id_ && start0 < last_end_))) { err = "runs must be sorted by chromosome and position"; return false; }
        for (size_t i = 0; i < n; i += ITEMS_PER_SLOT) {
            const uint32_t cnt = (uint32_t)std::min<size_t>(ITEMS_PER_SLOT, n - i);
            const uint32_t s = start0 + (uint32_t)i * step, e = s + (cnt - 1) * step + span;
            raw_.resize(24 + 4 * (size_t)cnt);
            const uint32_t h[5] = {id, s, e, step, span};
            memcpy(raw_.data(), h, 20);
            raw_[20] = 3;                                           // fixedStep
            raw_[21] = 0;
            const uint16_t c16 = (uint16_t)cnt;
            memcpy(raw_.data() + 22, &c16, 2);
            memcpy(raw_.data() + 24, v + i, 4 * (size_t)cnt);
            write_section(id, s, e);
            for (uint32_t k = 0; k < cnt; ++k) account(id, s + k * step, span, v[i + k]);
        }
        if (n) { last_id_ = id; last_end_ = start0 + (uint32_t)(n - 1) * step + span; }
        return true;
    }

    bool finish(std::string &err) {
        // zoom levels are built from level 0 summaries kept while the data was written (reduction 4x per level, from a base
        // resolution derived from the mean item span like wigToBigWig does)
        const uint64_t index_off = pos_;
        write_rtree(sections_);
        std::vector<ZoomHeader> zh;
        build_zooms(zh);
        const uint32_t max_raw = max_raw_;
        // patch the head
        std::vector<uint8_t> head(64 + 24 * MAX_ZOOM, 0);
        auto w16 = [&](size_t o, uint16_t v) { memcpy(head.data() + o, &v, 2); };
        auto w32 = [&](size_t o, uint32_t v) { memcpy(head.data() + o, &v, 4); };
        auto w64 = [&](size_t o, uint64_t v) { memcpy(head.data() + o, &v, 8); };
        w32(0, BW_MAGIC); w16(4, 4); w16(6, (uint16_t)zh.size()); w64(8, chrom_tree_off_); w64(16, data_off_); w64(24, index_off);
        w16(32, 0); w16(34, 0); w64(36, 0); w64(44, summary_off_); w32(52, compress_ ? max_raw : 0); w64(56, 0);
        for (size_t i = 0; i < zh.size(); ++i) { w32(64 + 24 * i, zh[i].reduction); w32(68 + 24 * i, 0); w64(72 + 24 * i, zh[i].data_off); w64(80 + 24 * i, zh[i].index_off); }
        uint8_t sum[40];
        const double mn = covered_ ? min_ : 0.0, mx = covered_ ? max_ : 0.0;
        memcpy(sum, &covered_, 8); memcpy(sum + 8, &mn, 8); memcpy(sum + 16, &mx, 8); memcpy(sum + 24, &sum_, 8); memcpy(sum + 32, &sumsq_, 8);
        const uint64_t n_sections = sections_.size();
        const uint32_t magic = BW_MAGIC;
        bool ok = fwrite(&magic, 4, 1, fo_) == 1;                                   // trailing magic
        ok = ok && fseeko(fo_, 0, SEEK_SET) == 0 && fwrite(head.data(), 1, head.size(), fo_) == head.size();
        ok = ok && fseeko(fo_, (off_t)summary_off_, SEEK_SET) == 0 && fwrite(sum, 1, 40, fo_) == 40;
        ok = ok && fseeko(fo_, (off_t)data_off_, SEEK_SET) == 0 && fwrite(&n_sections, 8, 1, fo_) == 1;
        ok = fclose(fo_) == 0 && ok;
        fo_ = nullptr;
        if (!ok) err = "write error";
        return ok;
    }

    ~BigWigWriter() { if (fo_) fclose(fo_); }

private:
    static constexpr int MAX_ZOOM = 10;
    struct Leaf { uint32_t c0, b0, c1, b1; uint64_t off, size; };
    struct ZoomHeader { uint32_t reduction; uint64_t data_off, index_off; };
    struct ZoomRec { uint32_t chrom, start, end, valid; float min, max, sum, sumsq; };

    FILE *fo_ = nullptr;
    bool compress_ = true;
    uint64_t pos_ = 0, summary_off_ = 0, chrom_tree_off_ = 0, data_off_ = 0;
    uint32_t key_size_ = 1, max_raw_ = 0, last_id_ = 0, last_end_ = 0;
    std::vector<std::pair<std::string, uint32_t>> chroms_;
    std::map<std::string, uint32_t> id_;
    std::vector<Leaf> sections_;
    std::vector<uint8_t> raw_, zbuf_;
    // statistics
    uint64_t covered_ = 0, span_total_ = 0, items_ = 0;
    double min_ = 0, max_ = 0, sum_ = 0, sumsq_ = 0;
    std::vector<ZoomRec> level0_;      // summaries at the first zoom resolution, built incrementally
    uint32_t reduction0_ = 0;

    void put(const std::vector<uint8_t> &v) { fwrite(v.data(), 1, v.size(), fo_); pos_ += v.size(); }
    void put(const void *p, size_t n) { fwrite(p, 1, n, fo_); pos_ += n; }
    void put64(uint64_t v) { put(&v, 8); }

    void write_chrom_tree() {
        // one leaf block when it fits, else a two-level tree (block size = number of chromosomes per node)
        const uint32_t n = (uint32_t)chroms_.size(), block = std::max<uint32_t>(1, std::min<uint32_t>(n, BLOCK_SIZE));
        const uint32_t hdr[4] = {BW_CHROM_TREE_MAGIC, block, key_size_, 8};
        put(hdr, 16);
        put64(n);
        put64(0);
        auto key = [&](const std::string &s) { std::vector<uint8_t> k(key_size_, 0); memcpy(k.data(), s.data(), s.size()); return k; };
        auto leaf_node = [&](uint32_t lo, uint32_t hi) {
            const uint8_t h[2] = {1, 0};
            const uint16_t c = (uint16_t)(hi - lo);
            put(h, 2); put(&c, 2);
            for (uint32_t i = lo; i < hi; ++i) { put(key(chroms_[i].first)); put(&i, 4); put(&chroms_[i].second, 4); }
        };
        if (n <= block) { leaf_node(0, n); return; }
        // levels: leaves of `block` items under index nodes of `block` children, as deep as needed
        struct Node { uint32_t lo, hi; };
        std::vector<std::vector<Node>> levels;
        std::vector<Node> cur;
        for (uint32_t i = 0; i < n; i += block) cur.push_back({i, std::min(n, i + block)});
        levels.push_back(cur);
        while (levels.back().size() > 1) {
            const auto &below = levels.back();
            std::vector<Node> up;
            for (size_t i = 0; i < below.size(); i += block) up.push_back({(uint32_t)i, (uint32_t)std::min(below.size(), i + block)});
            levels.push_back(up);
        }
        // sizes top-down to know the offsets
        const uint64_t leaf_item = key_size_ + 8, idx_item = key_size_ + 8;
        std::vector<uint64_t> level_off(levels.size());
        uint64_t off = pos_;
        for (int l = (int)levels.size() - 1; l >= 0; --l) {
            level_off[l] = off;
            for (auto &nd : levels[l]) off += 4 + (uint64_t)(nd.hi - nd.lo) * (l == 0 ? leaf_item : idx_item);
        }
        auto node_off = [&](int l, size_t idx) { uint64_t o = level_off[l]; for (size_t i = 0; i < idx; ++i) o += 4 + (uint64_t)(levels[l][i].hi - levels[l][i].lo) * (l == 0 ? leaf_item : idx_item); return o; };
        auto first_chrom = [&](int l, size_t idx) { size_t i = idx; for (int k = l; k > 0; --k) i = levels[k][i].lo; return levels[0][i].lo; };
        for (int l = (int)levels.size() - 1; l >= 1; --l)
            for (auto &nd : levels[l]) {
                const uint8_t h[2] = {0, 0};
                const uint16_t c = (uint16_t)(nd.hi - nd.lo);
                put(h, 2); put(&c, 2);
                for (uint32_t i = nd.lo; i < nd.hi; ++i) { put(key(chroms_[first_chrom(l - 1, i)].first)); put64(node_off(l - 1, i)); }
            }
        for (auto &nd : levels[0]) leaf_node(nd.lo, nd.hi);
    }

    void write_block(const std::vector<uint8_t> &raw, Leaf lf, std::vector<Leaf> &index) {
        max_raw_ = std::max<uint32_t>(max_raw_, (uint32_t)raw.size());
        lf.off = pos_;
        if (compress_) {
            uLongf zn = compressBound(raw.size());
            zbuf_.resize(zn);
            compress2(zbuf_.data(), &zn, raw.data(), raw.size(), 6);
            put(zbuf_.data(), zn);
            lf.size = zn;
        } else {
            put(raw.data(), raw.size());
            lf.size = raw.size();
        }
        index.push_back(lf);
    }

    void write_section(uint32_t id, uint32_t s, uint32_t e) { write_block(raw_, Leaf{id, s, id, e, 0, 0}, sections_); }

    void account(uint32_t id, uint32_t s, uint32_t span, float v) {
        if (!covered_) { min_ = max_ = v; }
        min_ = std::min<double>(min_, v);
        max_ = std::max<double>(max_, v);
        covered_ += span;
        sum_ += (double)v * span;
        sumsq_ += (double)v * v * span;
        span_total_ += span;
        ++items_;
        items0_.push_back(Item{id, s, s + span, v});
    }

    struct Item { uint32_t chrom, start, end; float v; };
    std::vector<Item> items0_;          // 16 bytes per item: kept in memory for the zoom pass (a 250 M-base track at step 3: 1.3 GB)

    // Summaries of `items` (sorted) at resolution `red`: bins are aligned to the first item they cover, as wigToBigWig does.
    static void reduce(const std::vector<ZoomRec> &in, uint32_t red, std::vector<ZoomRec> &out) {
        out.clear();
        for (const ZoomRec &r : in) {
            if (!out.empty() && out.back().chrom == r.chrom && r.end <= out.back().start + red) {
                ZoomRec &o = out.back();
                o.end = r.end; o.valid += r.valid; o.min = std::min(o.min, r.min); o.max = std::max(o.max, r.max); o.sum += r.sum; o.sumsq += r.sumsq;
            } else {
                out.push_back(r);
            }
        }
    }

    void build_zooms(std::vector<ZoomHeader> &zh) {
        if (items0_.empty()) return;
        // first reduction: 10 x the mean span, like bbiFile's initial zoom choice; then x4 per level while it shrinks the data
        uint32_t red = std::max<uint32_t>(10, (uint32_t)(10 * span_total_ / std::max<uint64_t>(1, items_)));
        std::vector<ZoomRec> cur, next;
        cur.reserve(items0_.size());
        for (const Item &it : items0_) cur.push_back(ZoomRec{it.chrom, it.start, it.end, it.end - it.start, it.v, it.v, it.v * (float)(it.end - it.start), it.v * it.v * (float)(it.end - it.start)});
        std::vector<Item>().swap(items0_);
        size_t prev = cur.size();
        for (int lvl = 0; lvl < MAX_ZOOM; ++lvl, red *= 4) {
            reduce(cur, red, next);
            if (next.size() * 2 > prev && lvl > 0) break;              // no longer shrinking
            cur.swap(next);
            prev = cur.size();
            ZoomHeader h{red, pos_, 0};
            const uint32_t nrec = (uint32_t)cur.size();
            put(&nrec, 4);
            std::vector<Leaf> index;
            std::vector<uint8_t> raw;
            for (size_t i = 0; i < cur.size(); i += ITEMS_PER_SLOT) {
                const size_t e = std::min(cur.size(), i + ITEMS_PER_SLOT);
                size_t j = i;
                while (j < e) {                                       // a block does not cross chromosomes
                    size_t k = j;
                    while (k < e && cur[k].chrom == cur[j].chrom) ++k;
                    raw.resize((k - j) * 32);
                    memcpy(raw.data(), &cur[j], raw.size());
                    write_block(raw, Leaf{cur[j].chrom, cur[j].start, cur[k - 1].chrom, cur[k - 1].end, 0, 0}, index);
                    j = k;
                }
            }
            h.index_off = pos_;
            write_rtree(index);
            zh.push_back(h);
            if (cur.size() <= 1) break;
        }
    }

    // R-tree over `leaves` (in file order): leaf nodes of BLOCK_SIZE items, index levels above them.
    void write_rtree(const std::vector<Leaf> &leaves) {
        const uint64_t n = leaves.size();
        uint32_t hdr[2] = {BW_RTREE_MAGIC, BLOCK_SIZE};
        put(hdr, 8);
        put64(n);
        const uint32_t bounds[4] = {n ? leaves.front().c0 : 0, n ? leaves.front().b0 : 0, n ? leaves.back().c1 : 0, n ? max_end(leaves, leaves.size() - 1, leaves.size()) : 0};
        put(bounds, 16);
        put64(n ? leaves.back().off + leaves.back().size : pos_);
        const uint32_t tail[2] = {ITEMS_PER_SLOT, 0};
        put(tail, 8);
        // levels[0] = groups of leaves; levels[k] = groups of level k-1 nodes
        struct Node { size_t lo, hi; uint32_t c0, b0, c1, b1; };
        std::vector<std::vector<Node>> levels;
        std::vector<Node> cur;
        for (size_t i = 0; i < n || (n == 0 && i == 0); i += BLOCK_SIZE) {
            const size_t hi = std::min<size_t>(n, i + BLOCK_SIZE);
            Node nd{i, hi, 0, 0, 0, 0};
            if (n) { nd.c0 = leaves[i].c0; nd.b0 = leaves[i].b0; nd.c1 = leaves[hi - 1].c1; nd.b1 = max_end(leaves, i, hi); }
            cur.push_back(nd);
            if (n == 0) break;
        }
        levels.push_back(cur);
        while (levels.back().size() > 1) {
            const auto &below = levels.back();
            std::vector<Node> up;
            for (size_t i = 0; i < below.size(); i += BLOCK_SIZE) {
                const size_t hi = std::min(below.size(), i + BLOCK_SIZE);
                Node nd{i, hi, below[i].c0, below[i].b0, below[hi - 1].c1, 0};
                for (size_t k = i; k < hi; ++k) if (below[k].c1 == nd.c1) nd.b1 = std::max(nd.b1, below[k].b1);
                up.push_back(nd);
            }
            levels.push_back(up);
        }
        std::vector<uint64_t> level_off(levels.size());
        uint64_t off = pos_;
        for (int l = (int)levels.size() - 1; l >= 0; --l) {
            level_off[l] = off;
            for (auto &nd : levels[l]) off += 4 + (uint64_t)(nd.hi - nd.lo) * (l == 0 ? 32 : 24);
        }
        auto node_off = [&](int l, size_t idx) { uint64_t o = level_off[l]; for (size_t i = 0; i < idx; ++i) o += 4 + (uint64_t)(levels[l][i].hi - levels[l][i].lo) * (l == 0 ? 32 : 24); return o; };
        for (int l = (int)levels.size() - 1; l >= 1; --l) {
            // offsets of the children accumulate; avoid the quadratic node_off for big levels
            uint64_t child = level_off[l - 1];
            size_t child_idx = 0;
            for (auto &nd : levels[l]) {
                const uint8_t h[2] = {0, 0};
                const uint16_t c = (uint16_t)(nd.hi - nd.lo);
                put(h, 2); put(&c, 2);
                for (size_t i = nd.lo; i < nd.hi; ++i) {
                    while (child_idx < i) { child += 4 + (uint64_t)(levels[l - 1][child_idx].hi - levels[l - 1][child_idx].lo) * (l - 1 == 0 ? 32 : 24); ++child_idx; }
                    const Node &b = levels[l - 1][i];
                    const uint32_t r[4] = {b.c0, b.b0, b.c1, b.b1};
                    put(r, 16);
                    put64(child);
                }
            }
        }
        (void)node_off;
        for (auto &nd : levels[0]) {
            const uint8_t h[2] = {1, 0};
            const uint16_t c = (uint16_t)(nd.hi - nd.lo);
            put(h, 2); put(&c, 2);
            for (size_t i = nd.lo; i < nd.hi; ++i) { const uint32_t r[4] = {leaves[i].c0, leaves[i].b0, leaves[i].c1, leaves[i].b1}; put(r, 16); put64(leaves[i].off); put64(leaves[i].size); }
        }
    }

    static uint32_t max_end(const std::vector<Leaf> &v, size_t lo, size_t hi) {
        uint32_t m = 0;
        const uint32_t c1 = v[hi - 1].c1;
        for (size_t i = lo; i < hi; ++i) if (v[i].c1 == c1) m = std::max(m, v[i].b1);
        return m;
    }
};

}  // namespace host
