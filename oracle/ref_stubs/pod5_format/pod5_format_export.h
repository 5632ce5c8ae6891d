/* TEST SCAFFOLDING for oracle/_ref only: cmake-generated export header of the pod5 submodule
 * (not generated here); the real pod5 c_api.h is only parsed, never linked. */
#pragma once
#define POD5_FORMAT_EXPORT
#define POD5_FORMAT_NO_EXPORT
#define POD5_FORMAT_DEPRECATED
