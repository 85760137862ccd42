// Host-side mirror of the reference's entry points, above the C-ABI (include/imageencoder_b200.h).
//
// Same class names, constructor arguments and process()/saveResult() protocol as the reference
// (ImageEncoder.hpp:16-22, ImageDecoder.hpp:15-19, VideoEncoder.hpp:13-15, VideoDecoder.hpp:15-17), so main() reads like
// main.cpp:105-150.  What the reference fixes at compile time is a run-time member here: block size (Block.hpp:13) is
// taken from the quant matrix, Huffman (makefile:13 -DENABLE_HUFFMAN) from dc::Options.  Errors are exceptions instead of
// exit()/assert() (ImageBase.cpp:24-27, ImageEncoder.cpp:26-28).  All arithmetic happens on the GPU through the C-ABI.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dc {

struct Options {
    bool huffman = true;      // the reference's default build has ENABLE_HUFFMAN (makefile:13)
    unsigned block = 0;       // 0: take it from the quant matrix (encoder) / 4 (decoder)
};

// code of a CodecError raised because an input file cannot be read: the reference exit(-1)s there (ImageBase.cpp:24-27,
// 103-106; VideoBase.cpp), which the CLIs mirror as exit status 255
constexpr int kUnreadableInput = -1000;

struct CodecError : std::runtime_error {
    int code;
    CodecError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// key=value settings file (ConfigReader.hpp:64-83; keys ConfigReader.cpp:18-38)
class ConfigReader {
    std::map<std::string, std::string> kv_;
    std::string err_;
  public:
    bool read(const std::string &file);
    bool verifyForImage();                       // exactly the 8 image keys (ConfigReader.cpp:185-207)
    bool verifyForVideo(bool encoder);           // >= 8 incl. gop, merange / >= 3 incl. motioncompensation (:209-242)
    std::string getValue(const std::string &key) const;
    std::string toString() const;
    const std::string &getErrorDescription() const { return err_; }
};

// N x N quant matrix from a text file, N = 4 or 8 decided by the file (MatrixReader.cpp:65-134)
// util::lexical_cast<uint16_t> of the reference (utils.hpp:293-305)
bool lexicalCastU16(const std::string &text, uint16_t &out);

class MatrixReader {
    std::vector<uint16_t> m_;
    unsigned n_ = 0;
  public:
    bool read(const std::string &file);
    const uint16_t *data() const { return m_.data(); }
    unsigned size() const { return n_; }
    std::string toString() const;
};

class ImageEncoder {
    std::string dest_;
    uint16_t w_, h_;
    bool rle_;
    MatrixReader q_;
    Options opt_;
    std::vector<uint8_t> raw_, out_;
  public:
    ImageEncoder(const std::string &source_file, const std::string &dest_file, const uint16_t &width, const uint16_t &height,
                 const bool &use_rle, MatrixReader &quant_m, const Options &opt = Options());
    bool process();
    void saveResult() const;
    const std::vector<uint8_t> &result() const { return out_; }
};

class ImageDecoder {
    std::string dest_;
    Options opt_;
    std::vector<uint8_t> enc_, out_;
    uint32_t w_ = 0, h_ = 0;
  public:
    ImageDecoder(const std::string &source_file, const std::string &dest_file, const Options &opt = Options());
    bool process();
    void saveResult() const;
    uint32_t width() const { return w_; }
    uint32_t height() const { return h_; }
};

class VideoEncoder {
    std::string dest_;
    uint16_t w_, h_, gop_, merange_;
    bool rle_;
    MatrixReader q_;
    Options opt_;
    std::vector<uint8_t> raw_, out_;
  public:
    VideoEncoder(const std::string &source_file, const std::string &dest_file, const uint16_t &width, const uint16_t &height,
                 const bool &use_rle, MatrixReader &quant_m, const uint16_t &gop, const uint16_t &merange,
                 const Options &opt = Options());
    bool process();
    void saveResult() const;
};

class VideoDecoder {
    std::string dest_;
    bool motioncomp_;
    std::vector<uint8_t> enc_, out_;
    uint32_t w_ = 0, h_ = 0, frames_ = 0;
  public:
    VideoDecoder(const std::string &source_file, const std::string &dest_file, const bool &motioncomp);
    bool process();
    void saveResult() const;
};

std::vector<uint8_t> readBinaryFile(const std::string &file);                       // throws CodecError
void writeBinaryFile(const std::string &file, const std::vector<uint8_t> &data);     // throws CodecError

}  // namespace dc
