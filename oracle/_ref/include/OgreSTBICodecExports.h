#ifndef OGRE_STBI_EXPORTS_H
#define OGRE_STBI_EXPORTS_H
#define _OgreSTBICodecExport __attribute__((visibility("default")))
#endif
