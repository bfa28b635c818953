// nimpress -- command-line front end; everything lives in libnimpress_host.so (nph_main).
#include "../../include/nimpress_host.h"
int main(int argc, char **argv) { return nph_main(argc, argv); }
