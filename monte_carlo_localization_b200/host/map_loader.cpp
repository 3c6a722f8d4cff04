// host/map_loader.cpp -- see map_loader.hpp.
#include "map_loader.hpp"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace particle_filter_cpp {

namespace {

std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}

std::string unquote(std::string s) {
    s = trim(s);
    if (s.size() >= 2 && (s.front() == '\'' || s.front() == '"') && s.back() == s.front()) s = s.substr(1, s.size() - 2);
    return s;
}

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

std::string dirname_of(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

// ---- PGM ---------------------------------------------------------------------------------
bool load_pgm(const std::vector<uint8_t>& buf, Image8& out, std::string* err) {
    size_t pos = 2;
    auto next_int = [&](long& v) -> bool {
        for (;;) {
            while (pos < buf.size() && std::isspace(buf[pos])) ++pos;
            if (pos < buf.size() && buf[pos] == '#') {
                while (pos < buf.size() && buf[pos] != '\n') ++pos;
                continue;
            }
            break;
        }
        if (pos >= buf.size() || !std::isdigit(buf[pos])) return false;
        v = 0;
        while (pos < buf.size() && std::isdigit(buf[pos])) v = v * 10 + (buf[pos++] - '0');
        return true;
    };
    const bool binary = buf[1] == '5';
    long w, h, maxv;
    if (!next_int(w) || !next_int(h) || !next_int(maxv) || w <= 0 || h <= 0 || maxv <= 0 || maxv > 65535) {
        if (err) *err = "bad PGM header";
        return false;
    }
    out.width = static_cast<int>(w);
    out.height = static_cast<int>(h);
    out.channels = 1;
    out.pix.resize(static_cast<size_t>(w) * h);
    if (binary) {
        ++pos;  // single whitespace after maxval
        const size_t bps = maxv > 255 ? 2 : 1;
        if (buf.size() - pos < static_cast<size_t>(w) * h * bps) {
            if (err) *err = "truncated PGM";
            return false;
        }
        for (size_t i = 0; i < out.pix.size(); ++i) {
            const long v = bps == 2 ? (buf[pos + 2 * i] << 8 | buf[pos + 2 * i + 1]) : buf[pos + i];
            out.pix[i] = static_cast<uint8_t>(v * 255 / maxv);
        }
    } else {
        for (size_t i = 0; i < out.pix.size(); ++i) {
            long v;
            if (!next_int(v)) {
                if (err) *err = "truncated ASCII PGM";
                return false;
            }
            out.pix[i] = static_cast<uint8_t>(v * 255 / maxv);
        }
    }
    return true;
}

// ---- PNG ---------------------------------------------------------------------------------
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

bool load_png(const std::vector<uint8_t>& buf, Image8& out, std::string* err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) {
        if (err) *err = "not a PNG";
        return false;
    }
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const char* type = reinterpret_cast<const char*>(&buf[pos + 4]);
        if (pos + 12 + len > buf.size()) break;
        const uint8_t* d = &buf[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            w = be32(d);
            h = be32(d + 4);
            depth = d[8];
            ctype = d[9];
            interlace = d[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(d, d + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (w == 0 || h == 0 || interlace != 0 || !(depth == 8 || (depth < 8 && (ctype == 0 || ctype == 3)))) {
        if (err) *err = "unsupported PNG (need non-interlaced, <= 8 bit)";
        return false;
    }
    int src_ch;
    switch (ctype) {
        case 0: src_ch = 1; break;
        case 2: src_ch = 3; break;
        case 3: src_ch = 1; break;
        case 4: src_ch = 2; break;
        case 6: src_ch = 4; break;
        default:
            if (err) *err = "bad PNG colour type";
            return false;
    }
    const size_t bpp_bits = static_cast<size_t>(src_ch) * depth;
    const size_t stride = (w * bpp_bits + 7) / 8;
    const size_t bpp = std::max<size_t>(1, bpp_bits / 8);
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf raw_len = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || raw_len != raw.size()) {
        if (err) *err = "PNG inflate failed";
        return false;
    }
    std::vector<uint8_t> img(stride * h);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t* in = &raw[y * (stride + 1) + 1];
        uint8_t* cur = &img[y * stride];
        const uint8_t* up = y ? &img[(y - 1) * stride] : nullptr;
        for (size_t x = 0; x < stride; ++x) {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
            int v = in[x];
            switch (ft) {
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) / 2; break;
                case 4: v += paeth(a, b, c); break;
                default: break;
            }
            cur[x] = static_cast<uint8_t>(v);
        }
    }
    auto sample = [&](uint32_t y, uint32_t x) -> int {   // sub-byte samples for depth < 8
        const uint8_t* row = &img[y * stride];
        if (depth == 8) return row[x];
        const int per = 8 / depth, shift = (per - 1 - static_cast<int>(x % per)) * depth;
        return (row[x / per] >> shift) & ((1 << depth) - 1);
    };
    out.width = static_cast<int>(w);
    out.height = static_cast<int>(h);
    if (ctype == 3) {
        const bool alpha = !trns.empty();
        out.channels = alpha ? 4 : 3;
        out.pix.resize(static_cast<size_t>(w) * h * out.channels);
        for (uint32_t y = 0; y < h; ++y)
            for (uint32_t x = 0; x < w; ++x) {
                const size_t idx = static_cast<size_t>(sample(y, x));
                uint8_t* o = &out.pix[(static_cast<size_t>(y) * w + x) * out.channels];
                for (int k = 0; k < 3; ++k) o[k] = 3 * idx + k < plte.size() ? plte[3 * idx + k] : 0;
                if (alpha) o[3] = idx < trns.size() ? trns[idx] : 255;
            }
    } else if (depth < 8) {   // low-depth gray: scale to 8 bits
        out.channels = 1;
        out.pix.resize(static_cast<size_t>(w) * h);
        const int maxv = (1 << depth) - 1;
        for (uint32_t y = 0; y < h; ++y)
            for (uint32_t x = 0; x < w; ++x) out.pix[static_cast<size_t>(y) * w + x] = static_cast<uint8_t>(sample(y, x) * 255 / maxv);
    } else {
        out.channels = src_ch;
        out.pix = img;
    }
    return true;
}

}  // namespace

bool parse_map_yaml(const std::string& path, MapYaml& out, std::string* err) {
    std::ifstream f(path);
    if (!f) {
        if (err) *err = "cannot open " + path;
        return false;
    }
    std::string line, list_key;
    int list_idx = 0;
    bool have_res = false, have_img = false;
    auto set = [&](const std::string& key, const std::string& val) {
        if (key == "image") {
            out.image = unquote(val);
            have_img = true;
        } else if (key == "resolution") {
            out.resolution = std::atof(val.c_str());
            have_res = true;
        } else if (key == "negate") {
            out.negate = std::atoi(val.c_str());
        } else if (key == "occupied_thresh") {
            out.occupied_thresh = std::atof(val.c_str());
        } else if (key == "free_thresh") {
            out.free_thresh = std::atof(val.c_str());
        }
    };
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line = line.substr(0, hash);
        std::string s = trim(line);
        if (s.empty()) continue;
        if (s[0] == '-' && !list_key.empty()) {   // block list item (origin:\n- x\n- y\n- yaw)
            if (list_key == "origin" && list_idx < 3) out.origin[list_idx++] = std::atof(trim(s.substr(1)).c_str());
            continue;
        }
        const size_t colon = s.find(':');
        if (colon == std::string::npos) continue;
        const std::string key = trim(s.substr(0, colon));
        std::string val = trim(s.substr(colon + 1));
        list_key.clear();
        if (val.empty()) {
            list_key = key;
            list_idx = 0;
            continue;
        }
        if (val[0] == '[') {   // flow list
            val = val.substr(1, val.find(']') == std::string::npos ? std::string::npos : val.find(']') - 1);
            std::stringstream ss(val);
            std::string tok;
            int k = 0;
            while (std::getline(ss, tok, ',') && k < 3) {
                if (key == "origin") out.origin[k] = std::atof(trim(tok).c_str());
                ++k;
            }
            continue;
        }
        set(key, val);
    }
    if (!have_res || !have_img) {
        if (err) *err = "map yaml needs image and resolution";
        return false;
    }
    return true;
}

bool load_image(const std::string& path, Image8& out, std::string* err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf) || buf.size() < 8) {
        if (err) *err = "cannot read " + path;
        return false;
    }
    if (buf[0] == 'P' && (buf[1] == '5' || buf[1] == '2')) return load_pgm(buf, out, err);
    return load_png(buf, out, err);
}

void image_to_grid(const Image8& img, const MapYaml& meta, OccupancyGrid& out) {
    const int w = img.width, h = img.height, ch = img.channels;
    out.width = static_cast<uint32_t>(w);
    out.height = static_cast<uint32_t>(h);
    out.data.assign(static_cast<size_t>(w) * h, -1);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const uint8_t* p = &img.pix[(static_cast<size_t>(y) * w + x) * ch];
            double sum;
            int n;
            if (ch == 1) {
                sum = 3.0 * p[0];
                n = 3;
            } else if (ch == 2) {   // gray + alpha: gray replicated into r, g, b; alpha averaged in
                sum = 3.0 * p[0] + p[1];
                n = 4;
            } else if (ch == 3) {
                sum = double(p[0]) + p[1] + p[2];
                n = 3;
            } else {
                sum = double(p[0]) + p[1] + p[2] + p[3];
                n = 4;
            }
            const double shade = (sum / n) / 255.0;
            const double occ = meta.negate ? shade : 1.0 - shade;
            int8_t cell = -1;
            if (occ > meta.occupied_thresh)
                cell = 100;
            else if (occ < meta.free_thresh)
                cell = 0;
            out.data[static_cast<size_t>(h - 1 - y) * w + x] = cell;   // row 0 = bottom
        }
    }
    out.resolution = static_cast<float>(meta.resolution);
    out.origin_x = meta.origin[0];
    out.origin_y = meta.origin[1];
    out.origin_yaw = meta.origin[2];
}

bool load_map(const std::string& yaml_path, OccupancyGrid& out, std::string* err) {
    MapYaml meta;
    if (!parse_map_yaml(yaml_path, meta, err)) return false;
    std::string img_path = meta.image;
    if (img_path.empty() || img_path[0] != '/') img_path = dirname_of(yaml_path) + "/" + img_path;
    Image8 img;
    if (!load_image(img_path, img, err)) return false;
    image_to_grid(img, meta, out);
    return true;
}

}  // namespace particle_filter_cpp
