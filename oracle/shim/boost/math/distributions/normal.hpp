// empty stand-in: src/consensus.h includes this header and uses nothing from it
