/* al.h -- stand-in for absent OpenAL (audio player, out of scope).
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). */
#ifndef ORACLE_STUB_AL_H
#define ORACLE_STUB_AL_H
typedef unsigned int ALuint;
typedef int ALint;
typedef int ALenum;
typedef int ALsizei;
typedef float ALfloat;
typedef void ALvoid;
#define AL_NO_ERROR 0
#define AL_FALSE 0
#define AL_INVALID_NAME 0xA001
#define AL_INVALID_ENUM 0xA002
#define AL_INVALID_VALUE 0xA003
#define AL_INVALID_OPERATION 0xA004
#define AL_OUT_OF_MEMORY 0xA005
#define AL_LOOPING 0x1007
#define AL_GAIN 0x100A
#define AL_SOURCE_STATE 0x1010
#define AL_PLAYING 0x1012
#define AL_BUFFERS_PROCESSED 0x1016
#define AL_FORMAT_MONO16 0x1101
ALenum alGetError(void);
void alGenSources(ALsizei n, ALuint *sources);
void alGenBuffers(ALsizei n, ALuint *buffers);
void alDeleteBuffers(ALsizei n, const ALuint *buffers);
void alBufferData(ALuint buffer, ALenum format, const ALvoid *data, ALsizei size, ALsizei freq);
void alSourcei(ALuint source, ALenum param, ALint value);
void alSourcef(ALuint source, ALenum param, ALfloat value);
void alGetSourcei(ALuint source, ALenum param, ALint *value);
void alGetSourceiv(ALuint source, ALenum param, ALint *values);
void alSourceQueueBuffers(ALuint source, ALsizei nb, const ALuint *buffers);
void alSourceUnqueueBuffers(ALuint source, ALsizei nb, ALuint *buffers);
void alSourcePlay(ALuint source);
#endif
