// glsl_keywords.hpp — included only inside a shader translation unit, after every real header.
// storage qualifiers / layout are meaningless once every interface variable is a namespace global
#define in
#define out
#define uniform
#define flat
#define layout(...)
#define main shader_main
#define discard do { gl_Discarded = true; return; } while (0)
