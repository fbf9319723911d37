// ORACLE — TEST INFRASTRUCTURE.  Stand-in for <gnuradio-4.0/reflection.hpp>: settings are set as plain members by
// oracle/ref_blocks.cpp, so the reflection lists expand to nothing.
#pragma once
#define ENABLE_REFLECTION(...)
#define ENABLE_REFLECTION_FOR_TEMPLATE(...)
#define ENABLE_REFLECTION_FOR_TEMPLATE_FULL(...)
