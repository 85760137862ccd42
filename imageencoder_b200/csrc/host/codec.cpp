#include "codec.hpp"

#include <algorithm>
#include <cctype>

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "../../../include/imageencoder_b200.h"

namespace dc {

static void check(int rc, const char *what) {
    if (rc != IE_OK) throw CodecError(rc, std::string(what) + ": " + ie_last_error());
}

std::vector<uint8_t> readBinaryFile(const std::string &file) {
    std::ifstream f(file, std::ios::binary | std::ios::ate);
    if (!f.good()) throw CodecError(kUnreadableInput, "cannot read file '" + file + "'");
    const std::streamsize n = f.tellg();
    std::vector<uint8_t> v((size_t)n);
    f.seekg(0);
    if (n > 0) f.read(reinterpret_cast<char *>(v.data()), n);
    return v;
}

void writeBinaryFile(const std::string &file, const std::vector<uint8_t> &data) {
    std::ofstream f(file, std::ios::binary);
    if (!f.good()) throw CodecError(IE_EINVAL, "cannot write file '" + file + "'");
    f.write(reinterpret_cast<const char *>(data.data()), (std::streamsize)data.size());
}

// ---- ConfigReader ------------------------------------------------------------------------------------------------
static const char *kImageKeys[] = {"rawfile", "encfile", "decfile", "rle", "quantfile", "width", "height", "logfile"};
static const char *kVideoEncKeys[] = {"rawfile", "encfile", "rle", "quantfile", "width", "height", "gop", "merange"};
static const char *kVideoDecKeys[] = {"encfile", "decfile", "motioncompensation"};

bool ConfigReader::read(const std::string &file) {
    kv_.clear();
    std::ifstream f(file);
    if (!f.good()) { err_ = "Can't open file"; return false; }
    std::string line;
    while (std::getline(f, line)) {
        while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
        if (line.empty()) continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) { err_ = "Can't find '=' in line"; return false; }
        const std::string key = line.substr(0, eq), val = line.substr(eq + 1);
        if (key.empty()) { err_ = "Detected an empty key"; return false; }
        if (kv_.count(key)) { err_ = "Key '" + key + "' was found more than once!"; return false; }
        kv_[key] = val;
    }
    return true;
}

template <size_t K> static bool have_all(const std::map<std::string, std::string> &kv, const char *(&keys)[K], std::string &err) {
    err.clear();
    for (const char *k : keys)
        if (!kv.count(k)) err += std::string("Key not found: '") + k + "'.\n";
    return err.empty();
}

bool ConfigReader::verifyForImage() {
    if (kv_.size() != 8) { err_ = "Too many or too few settings in file for image en/decoder!"; return false; }
    return have_all(kv_, kImageKeys, err_);
}

bool ConfigReader::verifyForVideo(bool encoder) {
    if (encoder) {
        if (kv_.size() < 8) { err_ = "Too many or too few settings in file for video encoder!"; return false; }
        return have_all(kv_, kVideoEncKeys, err_);
    }
    if (kv_.size() < 3) { err_ = "Too many or too few settings in file for video decoder!"; return false; }
    return have_all(kv_, kVideoDecKeys, err_);
}

std::string ConfigReader::getValue(const std::string &key) const {
    auto it = kv_.find(key);
    return it == kv_.end() ? std::string() : it->second;
}

std::string ConfigReader::toString() const {
    std::ostringstream o;
    for (const auto &p : kv_) o << std::setw(18) << p.first << " = " << p.second << "\n";
    return o.str();
}

// ---- MatrixReader ------------------------------------------------------------------------------------------------
// util::lexical_cast<uint16_t> of the reference (utils.hpp:293-305), which both the matrix entries (MatrixReader.cpp:104) and the
// numeric settings (main.cpp:91-98) go through: formatted stream extraction into a uint16_t -- hexadecimal if the text starts
// with 0x / 0X, leading whitespace and a sign accepted ("-4" wraps to 65532), anything after the number ignored ("8px" is 8),
// failure if there is no number or it does not fit.  Same libstdc++ call, so the same corner cases.
bool lexicalCastU16(const std::string &text, uint16_t &out) {
    std::stringstream cast;
    if (text.size() >= 2 && text[0] == '0' && (text[1] == 'x' || text[1] == 'X')) cast << std::hex << text;
    else cast << text;
    uint16_t v;
    if (!(cast >> v)) return false;
    out = v;
    return true;
}

// Text format as the reference reads it (MatrixReader.cpp:65-134): lines split at '\n', each trimmed of whitespace at both ends
// (so CRLF files work) and with runs of spaces collapsed, items separated by SINGLE SPACES ONLY (a tab is not a separator), every
// item through lexical_cast<uint16_t>; a blank line is a row without columns, i.e. an error, and so is a trailing blank line
// (one row too many).  The only difference: the reference's matrix size is the compile-time BlockSize, here it is the number of
// rows (4 or 8).
bool MatrixReader::read(const std::string &file) {
    std::ifstream f(file);
    if (!f.good()) { fprintf(stderr, "[MatrixReader] cannot read '%s'\n", file.c_str()); return false; }
    std::stringstream ss;
    ss << f.rdbuf();
    std::vector<std::vector<uint16_t>> rows;
    std::string line, item;
    while (std::getline(ss, line)) {
        line.erase(line.begin(), std::find_if(line.begin(), line.end(), [](int c) { return !std::isspace(c); }));          // utils.hpp:33-56
        line.erase(std::find_if(line.rbegin(), line.rend(), [](int c) { return !std::isspace(c); }).base(), line.end());
        line.erase(std::unique(line.begin(), line.end(), [](char l, char r) { return l == ' ' && r == ' '; }), line.end());   // utils.hpp:88-94
        std::vector<uint16_t> r;
        std::stringstream iss(line);
        while (std::getline(iss, item, ' ')) {
            uint16_t v;
            if (!lexicalCastU16(item, v)) { fprintf(stderr, "[MatrixReader] bad entry '%s'\n", item.c_str()); return false; }
            r.push_back(v);
        }
        rows.push_back(r);
    }
    const size_t n = rows.size();
    if (n != 4 && n != 8) { fprintf(stderr, "[MatrixReader] expected 4 or 8 rows, got %zu\n", n); return false; }
    m_.clear();
    for (const auto &r : rows) {
        if (r.size() != n) { fprintf(stderr, "[MatrixReader] expected %zu columns, got %zu\n", n, r.size()); return false; }
        for (uint16_t v : r) m_.push_back(v);
    }
    n_ = (unsigned)n;
    return true;
}

std::string MatrixReader::toString() const {
    std::ostringstream o;
    for (unsigned r = 0; r < n_; r++) {
        for (unsigned c = 0; c < n_; c++) o << std::setw(4) << m_[r * n_ + c];
        o << "\n";
    }
    return o.str();
}

// ---- images --------------------------------------------------------------------------------------------------------
ImageEncoder::ImageEncoder(const std::string &src, const std::string &dst, const uint16_t &w, const uint16_t &h, const bool &rle,
                           MatrixReader &q, const Options &opt)
    : dest_(dst), w_(w), h_(h), rle_(rle), q_(q), opt_(opt), raw_(readBinaryFile(src)) {
    const unsigned n = opt_.block ? opt_.block : q_.size();
    if (n != q_.size()) throw CodecError(IE_EINVAL, "quant matrix size does not match the block size");
    if (w_ % n || h_ % n) throw CodecError(IE_EINVAL, "width/height must be multiples of the block size");
    if (raw_.size() != (size_t)w_ * h_) throw CodecError(IE_EINVAL, "raw file size must be width*height");
    opt_.block = n;
}

bool ImageEncoder::process() {
    out_.resize(ie_max_encoded_bytes(w_, h_, opt_.block, 1));
    size_t n = 0;
    check(ie_encode_image(raw_.data(), w_, h_, opt_.block, q_.data(), rle_ ? 1 : 0, opt_.huffman ? 1 : 0, out_.data(), out_.size(), &n),
          "ImageEncoder::process");
    out_.resize(n);
    return true;
}

void ImageEncoder::saveResult() const {
    writeBinaryFile(dest_, out_);
    printf("[ImageProcessor] Original file size: %8zu bytes\n", raw_.size());
    printf("[ImageProcessor]       Encoded size: %8zu bytes  => Ratio: %.2f%%\n", out_.size(), 100.0f * float(out_.size()) / raw_.size());
    printf("[ImageProcessor] Saved file at: %s\n", dest_.c_str());
}

ImageDecoder::ImageDecoder(const std::string &src, const std::string &dst, const Options &opt)
    : dest_(dst), opt_(opt), enc_(readBinaryFile(src)) {
    if (!opt_.block) opt_.block = 4;
}

bool ImageDecoder::process() {
    out_.resize(16);
    int rc = ie_decode_image(enc_.data(), enc_.size(), opt_.block, out_.data(), out_.size(), &w_, &h_);
    if (rc == IE_ENOSPC && w_ && h_) {
        out_.resize((size_t)w_ * h_);
        rc = ie_decode_image(enc_.data(), enc_.size(), opt_.block, out_.data(), out_.size(), &w_, &h_);
    }
    check(rc, "ImageDecoder::process");
    out_.resize((size_t)w_ * h_);
    return true;
}

void ImageDecoder::saveResult() const {
    writeBinaryFile(dest_, out_);
    printf("[ImageProcessor] Original file size: %8zu bytes\n", enc_.size());
    printf("[ImageProcessor]       Decoded size: %8zu bytes  => Ratio: %.2f%%\n", out_.size(), 100.0f * float(out_.size()) / enc_.size());
    printf("[ImageProcessor] Saved file at: %s\n", dest_.c_str());
}

// ---- video ---------------------------------------------------------------------------------------------------------
VideoEncoder::VideoEncoder(const std::string &src, const std::string &dst, const uint16_t &w, const uint16_t &h, const bool &rle,
                           MatrixReader &q, const uint16_t &gop, const uint16_t &merange, const Options &opt)
    : dest_(dst), w_(w), h_(h), gop_(gop < 1 ? 1 : gop), merange_(merange), rle_(rle), q_(q), opt_(opt), raw_(readBinaryFile(src)) {
    if (q_.size() != 4) throw CodecError(IE_EINVAL, "video needs a 4x4 quant matrix (micro blocks are 4x4, ImageBase.cpp:266-306)");
    const size_t fsz = (size_t)w_ * h_ * 3 / 2;
    if (!fsz || raw_.size() % fsz) throw CodecError(IE_EINVAL, "file size must be a multiple of the YUV420 frame size");
}

bool VideoEncoder::process() {
    const size_t fsz = (size_t)w_ * h_ * 3 / 2;
    out_.resize(ie_max_encoded_bytes(w_, h_, 4, (uint32_t)std::max<size_t>(1, raw_.size() / fsz)));
    size_t n = 0;
    check(ie_encode_video(raw_.data(), raw_.size(), w_, h_, q_.data(), rle_ ? 1 : 0, gop_, merange_, opt_.huffman ? 1 : 0, out_.data(),
                          out_.size(), &n), "VideoEncoder::process");
    out_.resize(n);
    return true;
}

void VideoEncoder::saveResult() const {
    writeBinaryFile(dest_, out_);
    printf("[VideoProcessor] Original file size: %8zu bytes\n", raw_.size());
    printf("[VideoProcessor]       Encoded size: %8zu bytes  => Ratio: %.2f%%\n", out_.size(), 100.0f * float(out_.size()) / raw_.size());
    printf("[VideoProcessor] Saved file at: %s\n", dest_.c_str());
}

VideoDecoder::VideoDecoder(const std::string &src, const std::string &dst, const bool &mc)
    : dest_(dst), motioncomp_(mc), enc_(readBinaryFile(src)) {}

bool VideoDecoder::process() {
    out_.resize(16);
    size_t n = 0;
    int rc = ie_decode_video(enc_.data(), enc_.size(), motioncomp_ ? 1 : 0, out_.data(), out_.size(), &n, &w_, &h_, &frames_);
    if (rc == IE_ENOSPC && n) {
        out_.resize(n);
        rc = ie_decode_video(enc_.data(), enc_.size(), motioncomp_ ? 1 : 0, out_.data(), out_.size(), &n, &w_, &h_, &frames_);
    }
    check(rc, "VideoDecoder::process");
    out_.resize(n);
    return true;
}

void VideoDecoder::saveResult() const {
    writeBinaryFile(dest_, out_);
    printf("[VideoProcessor] Original file size: %8zu bytes\n", enc_.size());
    printf("[VideoProcessor]       Decoded size: %8zu bytes  => Ratio: %.2f%%\n", out_.size(), 100.0f * float(out_.size()) / enc_.size());
    printf("[VideoProcessor] Saved file at: %s\n", dest_.c_str());
}

}  // namespace dc
