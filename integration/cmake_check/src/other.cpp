// stand-in for the rest of vierkant's sources
int vierkant_other_stub() { return 0; }
