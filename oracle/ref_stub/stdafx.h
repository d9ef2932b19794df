// oracle/ref_stub/stdafx.h -- intentionally empty.
// Stands in for the reference's precompiled header (Source/stdafx.h:32 pulls atlbase.h) when
// oracle/Makefile compiles the reference's Source/csputils.cpp in place.
#pragma once
