#pragma once
namespace boost { struct progress_display { explicit progress_display(unsigned long) {} unsigned long operator+=(unsigned long x) { return c += x; } unsigned long operator++() { return ++c; } unsigned long c = 0; }; }
