// main.cpp — the reference's example driver (src/main.zig:8-43) on the CUDA engine:
// train on taylorswift.txt to vocab 300, write merges.txt, encode + decode the sample string, print ms.
#include <chrono>
#include <fstream>
#include <iterator>

#include "basic_tokenizer.hpp"

int main(int argc, char** argv) {
    const std::string text_path = argc > 1 ? argv[1] : "taylorswift.txt";
    const std::string merges_path = argc > 2 ? argv[2] : "merges.txt";
    const unsigned vocab = argc > 3 ? (unsigned)std::atoi(argv[3]) : 300;
    zigbpe::BasicTokenizer tokenizer;
    std::ifstream f(text_path, std::ios::binary);
    if (!f) { std::fprintf(stderr, "cannot open %s\n", text_path.c_str()); return 1; }
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto t0 = std::chrono::steady_clock::now();
    tokenizer.train(text, (uint16_t)vocab, false);
    tokenizer.serializeMerges(merges_path);
    auto tokens = tokenizer.encode("hello world!!!? (\xEC\x95\x88\xEB\x85\x95\xED\x95\x98\xEC\x84\xB8\xEC\x9A\x94!) lol123 \xF0\x9F\x98\x89");
    for (uint16_t t : tokens) std::fprintf(stderr, "%u ", (unsigned)t);
    std::string decoded = tokenizer.decode(tokens);
    std::fprintf(stderr, "\n%s\n", decoded.c_str());
    auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    std::fprintf(stderr, "Training completed in %lld ms\n", (long long)ms);
    return 0;
}
