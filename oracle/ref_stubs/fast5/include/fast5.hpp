/* TEST SCAFFOLDING for oracle/_ref only: the reference includes "../fast5/include/fast5.hpp"
 * (HDF5 wrapper, submodule not vendored here); nothing on the oracle's path uses it. */
#pragma once
