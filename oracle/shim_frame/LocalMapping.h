// ORACLE shim (test infrastructure): Frame.cc includes LocalMapping.h (src/Frame.cc:25) without using anything of it
#pragma once
