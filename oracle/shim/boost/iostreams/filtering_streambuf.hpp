// boost shim (see shim_core.hpp): test infrastructure only
#include <boost/iostreams/shim_core.hpp>
