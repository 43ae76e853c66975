// ORACLE BUILD STUB (test infrastructure): see Orthogonal_k_neighbor_search.h in this directory.
#include "Orthogonal_k_neighbor_search.h"
