// stand-in for <boost/optional.hpp> (TEST INFRASTRUCTURE): the reference's
// search_tree_cuda.cu includes the header but uses nothing from it.
#pragma once
