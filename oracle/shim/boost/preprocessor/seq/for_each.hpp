#pragma once
#include <boost/preprocessor/cat.hpp>
// shim: only supports the reference's usage BOOST_PP_SEQ_FOR_EACH(LOOP_BODY, _, SEQ)
#define BOOST_PP_SEQ_FOR_EACH(m, d, seq) BOOST_PP_CAT(SHIM_SFE_A seq, _END)
#define SHIM_SFE_A(x) LOOP_BODY(_, _, x) SHIM_SFE_B
#define SHIM_SFE_B(x) LOOP_BODY(_, _, x) SHIM_SFE_A
#define SHIM_SFE_A_END
#define SHIM_SFE_B_END
