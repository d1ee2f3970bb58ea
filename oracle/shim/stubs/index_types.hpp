// TEST INFRASTRUCTURE. Empty stand-in for the reference's include/index_types.hpp, on the include path of
// oracle/ref_drivers/ref_index.cpp only: include/ds2i/queries.hpp includes that header for its index aliases, but
// it needs real Boost.PP (absent here); the three DINT aliases are declared by the driver itself.
#pragma once
