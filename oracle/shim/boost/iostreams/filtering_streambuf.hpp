// empty stand-in: tracy's msa.h includes this header but the oracle uses nothing from it
