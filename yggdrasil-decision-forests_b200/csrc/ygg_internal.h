// ygg_internal.h — helpers shared by the translation units of libygg_b200.so (not part of the ABI).
#pragma once

// Records `msg` as this thread's ygg_last_error() and returns `code`.
__attribute__((visibility("hidden"))) int ygg_set_error_msg(int code, const char* msg);
