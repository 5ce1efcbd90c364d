// ref_curl_shim.cpp -- test infrastructure.  Thin extern "C" entry into the REFERENCE's own
// GenerateCurlNoise (ImageUtils.cpp:176-223), which is compiled verbatim next to this file by
// oracle/Makefile (target `ref`).  It writes a TGA; the caller decodes it.
#include <string>
void GenerateCurlNoise(std::string path);
extern "C" int ref_generate_curl_noise_tga(const char *path) {
    GenerateCurlNoise(std::string(path));
    return 0;
}
