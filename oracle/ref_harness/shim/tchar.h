/* Build shim for the reference's MSVC precompiled header (stdafx.h includes <tchar.h>).
 * Test infrastructure only; carries no codec arithmetic. */
#pragma once
typedef char _TCHAR;
