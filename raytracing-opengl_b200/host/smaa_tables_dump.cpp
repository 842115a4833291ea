/* smaa_tables_dump.cpp — writes the reference's SMAA lookup tables (src/AreaTex.h, src/SearchTex.h, included from where they
 * lie) as raw bytes into the asset mirror (host/build/assets/smaa/, git-ignored), next to the mirrored textures, so that hosts
 * which cannot include C headers (the Python mirror, bench.py) can hand them to rtb_smaa_set_tables().  Build-time tool. */
#include <cstdio>
#include <string>
#include <AreaTex.h>
#include <SearchTex.h>

static int dump(const std::string& path, const unsigned char* p, size_t n) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); return 1; }
    fwrite(p, 1, n, f);
    fclose(f);
    return 0;
}
int main(int argc, char** argv) {
    if (argc != 2) return 2;
    const std::string dir = argv[1];
    return dump(dir + "/area_rg8_160x560.bin", areaTexBytes, AREATEX_SIZE) | dump(dir + "/search_r8_64x16.bin", searchTexBytes, SEARCHTEX_SIZE);
}
