// gcov_harness.cpp -- measurement tooling (tools/gcov_ops.py): runs the reference's bc7enc_compress_block over a file of raw
// 4x4 RGBA tiles so that a --coverage build of the reference's bc7enc.cpp (compiled where it lies) yields per-line execution
// counts.  usage: gcov_harness <tiles.bin> <kind: all|opaque|alpha> <uber> <max_partitions> <filterbank 0|1>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bc7enc.h"

int main(int argc, char **argv)
{
    if(argc < 6) { return 2; }
    FILE *f = fopen(argv[1], "rb");
    if(!f) { return 3; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> tiles(static_cast<size_t>(bytes));
    if(fread(tiles.data(), 1, tiles.size(), f) != tiles.size()) { return 4; }
    fclose(f);
    const int kind = !strcmp(argv[2], "opaque") ? 1 : (!strcmp(argv[2], "alpha") ? 2 : 0);
    bc7enc_compress_block_params p;
    bc7enc_compress_block_params_init(&p);
    p.m_uber_level = uint32_t(atoi(argv[3]));
    p.m_max_partitions = uint32_t(atoi(argv[4]));
    p.m_mode17_partition_estimation_filterbank = atoi(argv[5]) != 0;
    bc7enc_compress_block_init();
    uint64_t n = 0, sum = 0;
    uint8_t out[16];
    for(size_t b = 0; b + 64 <= tiles.size(); b += 64)
    {
        bool alpha = false;
        for(int i = 0; i < 16; ++i) { alpha = alpha || tiles[b + 4 * i + 3] < 255; }// the dispatch of bc7enc_compress_block
        if((kind == 1 && alpha) || (kind == 2 && !alpha)) { continue; }
        bc7enc_compress_block(out, tiles.data() + b, &p);
        sum += out[0];
        ++n;
    }
    printf("%llu %llu\n", (unsigned long long) n, (unsigned long long) sum);
    return 0;
}
