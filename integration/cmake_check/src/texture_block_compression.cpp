// stand-in for the reference's src/texture_block_compression.cpp (the TU the module swaps out); never compiled with the option ON
int vierkant_cpu_bcn_stub() { return 0; }
