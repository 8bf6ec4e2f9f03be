/* alc.h -- stand-in for absent OpenAL context API.  TEST INFRASTRUCTURE ONLY. */
#ifndef ORACLE_STUB_ALC_H
#define ORACLE_STUB_ALC_H
typedef struct ALCdevice_struct ALCdevice;
typedef struct ALCcontext_struct ALCcontext;
typedef int ALCint;
typedef char ALCchar;
ALCdevice *alcOpenDevice(const ALCchar *name);
int alcCloseDevice(ALCdevice *device);
ALCcontext *alcCreateContext(ALCdevice *device, const ALCint *attrs);
int alcMakeContextCurrent(ALCcontext *ctx);
void alcDestroyContext(ALCcontext *ctx);
#endif
