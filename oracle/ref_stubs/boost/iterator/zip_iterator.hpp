// ORACLE BUILD STUB (test infrastructure): boost::make_zip_iterator / boost::tuple stand-ins live with the CGAL stub.
#include <CGAL/Orthogonal_k_neighbor_search.h>
