// fsea_lua.cpp -- headless script host: runs the reference's lua/fft*.lua scripts
// UNCHANGED against libfsea.so (the B200 engine behind the nrf_* / nut_* operator API).
//
// The reference's app shell (src/main.cpp) needs a window, OpenGL, OpenAL and OSC.  This
// host keeps what the scripts can observe of it: the same Lua globals, the same object
// representation -- a table {__type__ = "nrf_fft", __ptr__ = lightuserdata, ...} whose
// metatable carries the type's __gc (src/main.cpp:25-60,1095-1109) -- the same
// setup() / draw() / on_key(key, mods) protocol (API.md:3-12, src/main.cpp:1269,977-982),
// and a full garbage-collection cycle after every frame (src/main.cpp:1316).
//   real:   nut_buffer_*, nrf_device_*, nrf_block_connect, nrf_fft_*, nrf_freq_shifter_*
//   inert:  ngl_* (rendering), nwm_get_time, nosc_server_*, nrf_player_*  (out of scope)
//   tap:    ngl_texture_update(texture, buffer, w, h) -- the consumer of the spectrum;
//           narrows float64 -> float32 exactly as src/ngl.c:229-239 does and appends the
//           w*h floats to --dump FILE (and refuses w*h > buffer length like src/ngl.c:211-214).
//
// usage: fsea-lua SCRIPT [--frames K] [--dump FILE] [--step] [--key FRAME:KEY:MODS]...
//   --step   deterministic replay: the device is paused and advanced one capture block per
//            frame by this host (nrf_device_set_paused / nrf_device_step, src/nrf.c:341-350)
// Run with the working directory one level below a directory holding lua/ and rfdata/
// (the scripts say "../rfdata/..."; _keys.lua is loaded from "../lua/", src/main.cpp:1201).
extern "C" {
#include <lauxlib.h>
#include <lua.h>
#include <lualib.h>
}
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "nrf.h"
#include "nut.h"

extern "C" void nrf_device_publish_now(nrf_device* device);   // libfsea extension (deterministic replay)

static FILE* g_dump = nullptr;
static long g_updates = 0;
static double g_checksum = 0.0;
static int g_frame = 0;

// ---------------------------------------------------------------- object tables
static void push_object(lua_State* L, const char* type, void* ptr) {
    lua_newtable(L);
    luaL_getmetatable(L, type);
    lua_setmetatable(L, -2);
    lua_pushstring(L, type);
    lua_setfield(L, -2, "__type__");
    lua_pushlightuserdata(L, ptr);
    lua_setfield(L, -2, "__ptr__");
}

static void* check_object(lua_State* L, int index, const char* type) {
    luaL_checktype(L, index, LUA_TTABLE);
    lua_getfield(L, index, "__type__");
    const char* have = lua_tostring(L, -1);
    if (!have || strcmp(have, type) != 0) {            // fail-stop, as src/main.cpp:56-59
        fprintf(stderr, "Lua: invalid type for param %d: expected %s, was %s\n", index, type, have ? have : "(none)");
        exit(EXIT_FAILURE);
    }
    lua_pop(L, 1);
    lua_getfield(L, index, "__ptr__");
    void* p = lua_touserdata(L, -1);
    lua_pop(L, 1);
    return p;
}

static void register_type(lua_State* L, const char* type, lua_CFunction gc) {
    luaL_newmetatable(L, type);
    if (gc) {
        lua_pushcfunction(L, gc);
        lua_setfield(L, -2, "__gc");
    }
    lua_pop(L, 1);
}

// ---------------------------------------------------------------- nut_buffer
static int push_buffer(lua_State* L, nut_buffer* b) {
    push_object(L, "nut_buffer", b);
    lua_pushinteger(L, b->length); lua_setfield(L, -2, "length");
    lua_pushinteger(L, b->channels); lua_setfield(L, -2, "channels");
    lua_pushinteger(L, b->size_bytes); lua_setfield(L, -2, "size_bytes");
    return 1;
}
static nut_buffer* to_buffer(lua_State* L, int i) { return (nut_buffer*)check_object(L, i, "nut_buffer"); }
static int l_buffer_gc(lua_State* L) { nut_buffer_free(to_buffer(L, 1)); return 0; }
static int l_buffer_append(lua_State* L) { nut_buffer_append(to_buffer(L, 1), to_buffer(L, 2)); return 0; }
static int l_buffer_reduce(lua_State* L) { return push_buffer(L, nut_buffer_reduce(to_buffer(L, 1), luaL_checknumber(L, 2))); }
static int l_buffer_clip(lua_State* L) {
    return push_buffer(L, nut_buffer_clip(to_buffer(L, 1), (int)luaL_checkinteger(L, 2), (int)luaL_checkinteger(L, 3)));
}
static int l_buffer_convert(lua_State* L) {
    return push_buffer(L, nut_buffer_convert(to_buffer(L, 1), (nut_buffer_type)luaL_checkinteger(L, 2)));
}
static int l_buffer_save(lua_State* L) { nut_buffer_save(to_buffer(L, 1), lua_tostring(L, 2)); return 0; }

// ---------------------------------------------------------------- nrf_device
static nrf_device* to_device(lua_State* L, int i) { return (nrf_device*)check_object(L, i, "nrf_device"); }
static int push_device(lua_State* L, nrf_device* d) {
    push_object(L, "nrf_device", d);
    lua_pushinteger(L, d->sample_rate);
    lua_setfield(L, -2, "sample_rate");
    return 1;
}
static int l_device_new(lua_State* L) { return push_device(L, nrf_device_new(luaL_checknumber(L, 1), lua_tostring(L, 2))); }
static int l_device_new_with_config(lua_State* L) {
    nrf_device_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    if (lua_istable(L, 1)) {
        lua_getfield(L, 1, "sample_rate"); if (lua_isnumber(L, -1)) cfg.sample_rate = (int)lua_tointeger(L, -1); lua_pop(L, 1);
        lua_getfield(L, 1, "freq_mhz"); if (lua_isnumber(L, -1)) cfg.freq_mhz = lua_tonumber(L, -1); lua_pop(L, 1);
        lua_getfield(L, 1, "data_file"); if (lua_isstring(L, -1)) cfg.data_file = lua_tostring(L, -1); lua_pop(L, 1);
    }
    return push_device(L, nrf_device_new_with_config(cfg));
}
static int l_device_gc(lua_State* L) { nrf_device_free(to_device(L, 1)); return 0; }
static int l_device_set_frequency(lua_State* L) {
    lua_pushnumber(L, nrf_device_set_frequency(to_device(L, 1), luaL_checknumber(L, 2)));
    return 1;
}
static int l_device_set_paused(lua_State* L) {
    if (lua_isboolean(L, 2)) nrf_device_set_paused(to_device(L, 1), lua_toboolean(L, 2));
    else fprintf(stderr, "nrf_device_set_paused expects boolean as second argument.\n");
    return 0;
}
static int l_device_step(lua_State* L) { nrf_device_step(to_device(L, 1)); return 0; }
static int l_device_get_samples(lua_State* L) { return push_buffer(L, nrf_device_get_samples_buffer(to_device(L, 1))); }

// base-type access for nrf_block_connect: any block object's __ptr__ (src/main.cpp:616-631)
static nrf_block* to_block(lua_State* L, int i) {
    luaL_checktype(L, i, LUA_TTABLE);
    lua_getfield(L, i, "__ptr__");
    nrf_block* b = (nrf_block*)lua_touserdata(L, -1);
    lua_pop(L, 1);
    if (!b || b->type < NRF_BLOCK_SOURCE || b->type > NRF_BLOCK_SINK) {
        fprintf(stderr, "Lua: param %d is not an nrf block\n", i);
        exit(EXIT_FAILURE);
    }
    return b;
}
static int l_block_connect(lua_State* L) { nrf_block_connect(to_block(L, 1), to_block(L, 2)); return 0; }

// ---------------------------------------------------------------- nrf_fft / shifter
static nrf_fft* to_fft(lua_State* L, int i) { return (nrf_fft*)check_object(L, i, "nrf_fft"); }
static int l_fft_new(lua_State* L) {
    push_object(L, "nrf_fft", nrf_fft_new((int)luaL_checkinteger(L, 1), (int)luaL_checkinteger(L, 2)));
    return 1;
}
static int l_fft_gc(lua_State* L) { nrf_fft_free(to_fft(L, 1)); return 0; }
static int l_fft_shift(lua_State* L) {
    float d = (float)luaL_checknumber(L, 2);            // narrowed through float, src/main.cpp:789
    nrf_fft_shift(to_fft(L, 1), d);
    return 0;
}
static int l_fft_process(lua_State* L) { nrf_fft_process(to_fft(L, 1), to_buffer(L, 2)); return 0; }
static int l_fft_get_buffer(lua_State* L) { return push_buffer(L, nrf_fft_get_buffer(to_fft(L, 1))); }

static nrf_freq_shifter* to_shifter(lua_State* L, int i) { return (nrf_freq_shifter*)check_object(L, i, "nrf_freq_shifter"); }
static int l_shifter_new(lua_State* L) {
    // scripts pass float-valued integers (shift = -10e3): accept any number
    push_object(L, "nrf_freq_shifter", nrf_freq_shifter_new((int)luaL_checknumber(L, 1), (int)luaL_checknumber(L, 2)));
    return 1;
}
static int l_shifter_gc(lua_State* L) { nrf_freq_shifter_free(to_shifter(L, 1)); return 0; }
static int l_shifter_process(lua_State* L) { nrf_freq_shifter_process(to_shifter(L, 1), to_buffer(L, 2)); return 0; }
static int l_shifter_get_buffer(lua_State* L) { return push_buffer(L, nrf_freq_shifter_get_buffer(to_shifter(L, 1))); }

// nrf_iq_filter_* (src/main.cpp:719-747: new(sample_rate, half_ampl_freq, kernel_length), process(filter, buffer), get_buffer)
static nrf_iq_filter* to_iq_filter(lua_State* L, int i) { return (nrf_iq_filter*)check_object(L, i, "nrf_iq_filter"); }
static int l_iq_filter_new(lua_State* L) {
    push_object(L, "nrf_iq_filter", nrf_iq_filter_new((int)luaL_checknumber(L, 1), (int)luaL_checknumber(L, 2), (int)luaL_checknumber(L, 3)));
    return 1;
}
static int l_iq_filter_gc(lua_State* L) { nrf_iq_filter_free(to_iq_filter(L, 1)); return 0; }
static int l_iq_filter_process(lua_State* L) { nrf_iq_filter_process(to_iq_filter(L, 1), to_buffer(L, 2)); return 0; }
static int l_iq_filter_get_buffer(lua_State* L) { return push_buffer(L, nrf_iq_filter_get_buffer(to_iq_filter(L, 1))); }

// ---------------------------------------------------------------- inert rendering / audio / osc
static int l_nop(lua_State*) { return 0; }
template <int K> static int l_dummy(lua_State* L) {
    static const char* names[] = {"ngl_camera", "ngl_model", "ngl_shader", "ngl_texture", "ngl_skybox", "ngl_font", "nosc_server", "nrf_player"};
    static int cells[8];
    push_object(L, names[K], &cells[K]);
    return 1;
}
static int l_get_time(lua_State* L) { lua_pushnumber(L, g_frame / 60.0); return 1; }
static int l_osc_update(lua_State* L) { lua_newtable(L); return 1; }

// the consumer of the spectrum (src/ngl.c:204-241): float64 -> float32, width*height values
static int l_texture_update(lua_State* L) {
    (void)check_object(L, 1, "ngl_texture");
    nut_buffer* b = to_buffer(L, 2);
    const long w = (long)luaL_checkinteger(L, 3), h = (long)luaL_checkinteger(L, 4);
    const long count = w * h;
    if (count > (long)b->length * b->channels) {
        fprintf(stderr, "ngl_texture_update: buffer too small for %ldx%ld (length %d)\n", w, h, b->length);
        exit(EXIT_FAILURE);
    }
    std::vector<float> tex((size_t)count);
    for (long i = 0; i < count; i++) {
        tex[i] = (b->type == NUT_BUFFER_F64) ? (float)b->data.f64[i] : (float)b->data.u8[i];
        g_checksum += tex[i];
    }
    if (g_dump) {
        int32_t hdr[3] = {(int32_t)g_frame, (int32_t)w, (int32_t)h};
        fwrite(hdr, sizeof(hdr), 1, g_dump);
        fwrite(tex.data(), sizeof(float), (size_t)count, g_dump);
    }
    g_updates++;
    return 0;
}

// ---------------------------------------------------------------- host
static void reg(lua_State* L, const char* name, lua_CFunction fn) { lua_pushcfunction(L, fn); lua_setglobal(L, name); }
static void constant(lua_State* L, const char* name, int v) { lua_pushinteger(L, v); lua_setglobal(L, name); }

static bool call_global(lua_State* L, const char* name, int nargs_pushed = 0) {
    lua_getglobal(L, name);
    if (!lua_isfunction(L, -1)) { lua_pop(L, 1 + nargs_pushed); return false; }
    if (nargs_pushed) lua_insert(L, -1 - nargs_pushed);
    if (lua_pcall(L, nargs_pushed, 0, 0)) {
        fprintf(stderr, "Error calling %s(): %s\n", name, lua_tostring(L, -1));
        lua_pop(L, 1);
        return false;
    }
    return true;
}

struct KeyEvent { int frame; std::string key; int mods; };

int main(int argc, char** argv) {
    const char* script = nullptr;
    const char* dump = nullptr;
    const char* keys_file = "../lua/_keys.lua";
    int frames = 4;
    bool step = false;
    std::vector<KeyEvent> keys;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--frames") && i + 1 < argc) frames = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
        else if (!strcmp(argv[i], "--keys-file") && i + 1 < argc) keys_file = argv[++i];
        else if (!strcmp(argv[i], "--step")) step = true;
        else if (!strcmp(argv[i], "--key") && i + 1 < argc) {
            KeyEvent e; char name[64] = ""; e.mods = 0;
            if (sscanf(argv[++i], "%d:%63[^:]:%d", &e.frame, name, &e.mods) >= 2) { e.key = name; keys.push_back(e); }
        } else script = argv[i];
    }
    if (!script) { fprintf(stderr, "usage: fsea-lua SCRIPT [--frames K] [--dump FILE] [--step] [--key FRAME:KEY_NAME:MODS]\n"); return 2; }
    if (dump && !(g_dump = fopen(dump, "wb"))) { perror(dump); return 2; }

    lua_State* L = luaL_newstate();
    luaL_openlibs(L);
    register_type(L, "nut_buffer", l_buffer_gc);
    register_type(L, "nrf_device", l_device_gc);
    register_type(L, "nrf_fft", l_fft_gc);
    register_type(L, "nrf_freq_shifter", l_shifter_gc);
    register_type(L, "nrf_iq_filter", l_iq_filter_gc);
    for (const char* t : {"ngl_camera", "ngl_model", "ngl_shader", "ngl_texture", "ngl_skybox", "ngl_font", "nosc_server", "nrf_player"})
        register_type(L, t, nullptr);

    reg(L, "nut_buffer_append", l_buffer_append);
    reg(L, "nut_buffer_reduce", l_buffer_reduce);
    reg(L, "nut_buffer_clip", l_buffer_clip);
    reg(L, "nut_buffer_convert", l_buffer_convert);
    reg(L, "nut_buffer_save", l_buffer_save);
    reg(L, "nrf_block_connect", l_block_connect);
    reg(L, "nrf_device_new", l_device_new);
    reg(L, "nrf_device_new_with_config", l_device_new_with_config);
    reg(L, "nrf_device_set_frequency", l_device_set_frequency);
    reg(L, "nrf_device_set_paused", l_device_set_paused);
    reg(L, "nrf_device_step", l_device_step);
    reg(L, "nrf_device_get_samples_buffer", l_device_get_samples);
    reg(L, "nrf_fft_new", l_fft_new);
    reg(L, "nrf_fft_shift", l_fft_shift);
    reg(L, "nrf_fft_process", l_fft_process);
    reg(L, "nrf_fft_get_buffer", l_fft_get_buffer);
    reg(L, "nrf_freq_shifter_new", l_shifter_new);
    reg(L, "nrf_freq_shifter_process", l_shifter_process);
    reg(L, "nrf_freq_shifter_get_buffer", l_shifter_get_buffer);
    reg(L, "nrf_iq_filter_new", l_iq_filter_new);
    reg(L, "nrf_iq_filter_process", l_iq_filter_process);
    reg(L, "nrf_iq_filter_get_buffer", l_iq_filter_get_buffer);
    // out of scope: present so that the scripts run, inert
    reg(L, "ngl_camera_new", l_dummy<0>);
    reg(L, "ngl_camera_new_look_at", l_dummy<0>);
    reg(L, "ngl_model_new_grid_triangles", l_dummy<1>);
    reg(L, "ngl_model_new_grid_points", l_dummy<1>);
    reg(L, "ngl_model_new_with_height_map", l_dummy<1>);
    reg(L, "ngl_model_load_obj", l_dummy<1>);
    reg(L, "ngl_shader_new", l_dummy<2>);
    reg(L, "ngl_shader_new_from_file", l_dummy<2>);
    reg(L, "ngl_texture_new", l_dummy<3>);
    reg(L, "ngl_skybox_new", l_dummy<4>);
    reg(L, "ngl_font_new", l_dummy<5>);
    reg(L, "nosc_server_new", l_dummy<6>);
    reg(L, "nrf_player_new", l_dummy<7>);
    for (const char* n : {"ngl_clear", "ngl_clear_depth", "ngl_draw_model", "ngl_draw_background", "ngl_model_translate",
                          "ngl_camera_translate", "ngl_camera_rotate_x", "ngl_camera_rotate_y", "ngl_camera_rotate_z",
                          "ngl_shader_uniform_set_float", "ngl_font_draw", "ngl_capture_model",
                          "nrf_player_set_gain", "nrf_player_set_freq_offset"})
        reg(L, n, l_nop);
    reg(L, "ngl_texture_update", l_texture_update);
    reg(L, "nwm_get_time", l_get_time);
    reg(L, "nosc_server_update", l_osc_update);
    constant(L, "NUT_BUFFER_U8", NUT_BUFFER_U8);
    constant(L, "NUT_BUFFER_F64", NUT_BUFFER_F64);
    constant(L, "NRF_SAMPLES_LENGTH", NRF_SAMPLES_LENGTH);
    constant(L, "NRF_DEMODULATE_RAW", 0);
    constant(L, "NRF_DEMODULATE_WBFM", 1);
    constant(L, "GL_POINTS", 0); constant(L, "GL_LINES", 1); constant(L, "GL_LINE_LOOP", 2); constant(L, "GL_LINE_STRIP", 3);
    constant(L, "GL_TRIANGLES", 4); constant(L, "GL_TRIANGLE_STRIP", 5); constant(L, "GL_TRIANGLE_FAN", 6);

    if (luaL_loadfile(L, keys_file) || lua_pcall(L, 0, 0, 0)) {
        fprintf(stderr, "%s\n", lua_tostring(L, -1));
        lua_pop(L, 1);
    }
    if (luaL_loadfile(L, script) || lua_pcall(L, 0, 0, 0)) {
        fprintf(stderr, "%s\n", lua_tostring(L, -1));
        return 1;
    }
    if (!call_global(L, "setup")) return 1;

    nrf_device* dev = nullptr;
    if (step) {                                            // scripts keep their device in the global `device`
        lua_getglobal(L, "device");
        if (lua_istable(L, -1)) {
            lua_getfield(L, -1, "__ptr__");
            dev = (nrf_device*)lua_touserdata(L, -1);
            lua_pop(L, 1);
        }
        lua_pop(L, 1);
        if (dev) { nrf_device_set_paused(dev, 1); dev->dummy_block_index = 0; nrf_device_publish_now(dev); }
    }
    for (g_frame = 0; g_frame < frames; g_frame++) {
        for (const KeyEvent& e : keys)
            if (e.frame == g_frame) {
                lua_getglobal(L, e.key.c_str());
                lua_Integer code = lua_tointeger(L, -1);
                lua_pop(L, 1);
                lua_pushinteger(L, code);
                lua_pushinteger(L, e.mods);
                call_global(L, "on_key", 2);
            }
        if (!call_global(L, "draw")) return 1;
        lua_gc(L, LUA_GCCOLLECT, 0);                       // src/main.cpp:1316
        if (dev) { nrf_device_step(dev); nrf_device_publish_now(dev); }
    }
    lua_close(L);                                          // runs every __gc (hot-reload path, src/main.cpp:1281)
    if (g_dump) fclose(g_dump);
    printf("fsea-lua: %s: %d frames, %ld texture updates, checksum %.9g\n", script, frames, g_updates, g_checksum);
    return 0;
}
