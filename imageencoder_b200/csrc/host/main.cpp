// `encoder <conf>` / `decoder <conf>`: drop-in for the reference CLIs (main.cpp:19-185), same .conf keys, same exit codes
// (1 usage, 2 unreadable conf, 3 bad settings, 4 bad quant matrix, 5 bad number).  Built twice: -DENCODER / -DDECODER.
// Extra, optional flags in front of the conf (the reference decides these at compile time):
//   --huffman | --no-huffman   (default --huffman, as the reference makefile builds)      --block 4|8 (decoder; default 4)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "codec.hpp"

// numeric settings as main.cpp:91-98 reads them (util::lexical_cast<uint16_t>)
static bool to_u16(const std::string &s, uint16_t &out) { return dc::lexicalCastU16(s, out); }

int main(int argc, char **argv) {
    dc::Options opt;
    int ai = 1;
    for (; ai < argc && argv[ai][0] == '-' && argv[ai][1] == '-'; ai++) {
        if (!strcmp(argv[ai], "--huffman")) opt.huffman = true;
        else if (!strcmp(argv[ai], "--no-huffman")) opt.huffman = false;
        else if (!strcmp(argv[ai], "--block") && ai + 1 < argc) opt.block = (unsigned)atoi(argv[++ai]);
        else { std::cerr << "unknown option " << argv[ai] << std::endl; return 1; }
    }
    if (argc - ai != 1) {
        std::cerr << "One argument, the name of a settings file, expected!" << std::endl;
        return 1;
    }
    dc::ConfigReader c;
    if (!c.read(argv[ai])) {
        std::cerr << "Error reading file '" << argv[ai] << "'!" << std::endl << c.getErrorDescription() << std::endl;
        return 2;
    }
    const bool is_image = c.verifyForImage();
    const std::string e_img = c.getErrorDescription();
    const bool is_encvideo = c.verifyForVideo(true);
    const std::string e_ev = c.getErrorDescription();
    const bool is_decvideo = c.verifyForVideo(false);
    const std::string e_dv = c.getErrorDescription();
    if (!((is_image && !(is_encvideo || is_decvideo)) || ((is_encvideo || is_decvideo) && !is_image))) {
        std::cerr << "Error in settings!" << std::endl << e_img << std::endl << e_ev << std::endl << e_dv << std::endl;
        return 3;
    }
    std::cout << "Input settings:\n-------------------------\n" << c.toString() << std::endl;
    const std::string encfile = c.getValue("encfile"), decfile = c.getValue("decfile");
    auto t0 = std::chrono::steady_clock::now();
    auto elapsed = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    try {
#ifdef ENCODER
        const std::string rawfile = c.getValue("rawfile");
        if (rawfile == encfile) { std::cerr << "Error in settings! Encoded filename must be different from raw filename!" << std::endl; return 3; }
        dc::MatrixReader m;
        if (!m.read(c.getValue("quantfile"))) return 4;
        std::cout << "Quantization matrix:\n-------------------------\n" << m.toString() << std::endl;
        uint16_t width, height, rle, gop = 0, merange = 0;
        if (!to_u16(c.getValue("width"), width) || !to_u16(c.getValue("height"), height) || !to_u16(c.getValue("rle"), rle)) return 5;
        if (is_encvideo && (!to_u16(c.getValue("gop"), gop) || !to_u16(c.getValue("merange"), merange))) return 5;
        if (is_image) {
            dc::ImageEncoder enc(rawfile, encfile, width, height, rle != 0, m, opt);
            if (enc.process()) { enc.saveResult(); printf("\nElapsed time: %f milliseconds\n\n", elapsed()); }
        } else if (is_encvideo) {
            dc::VideoEncoder enc(rawfile, encfile, width, height, rle != 0, m, gop, merange, opt);
            if (enc.process()) { enc.saveResult(); printf("\nElapsed time: %f milliseconds\n\n", elapsed()); }
        }
#endif
#ifdef DECODER
        if (encfile == decfile) { std::cerr << "Error in settings! Decoded filename must be different from encoded!" << std::endl; return 3; }
        t0 = std::chrono::steady_clock::now();
        if (is_image) {
            dc::ImageDecoder dec(encfile, decfile, opt);
            if (dec.process()) { dec.saveResult(); printf("\nElapsed time: %f milliseconds\n", elapsed()); }
        } else if (is_decvideo) {
            uint16_t mc;
            if (!to_u16(c.getValue("motioncompensation"), mc)) return 5;
            dc::VideoDecoder dec(encfile, decfile, mc != 0);
            if (dec.process()) { dec.saveResult(); printf("\nElapsed time: %f milliseconds\n", elapsed()); }
        }
#endif
    } catch (const dc::CodecError &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return e.code == dc::kUnreadableInput ? 255 : 6;      // 255: the reference's exit(-1) on an unreadable input file
    }
    return 0;
}
