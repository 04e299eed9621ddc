#ifndef OGRE_EXPORTS_H
#define OGRE_EXPORTS_H
#define _OgreExport __attribute__((visibility("default")))
#define _OgrePrivate __attribute__((visibility("hidden")))
#define OGRE_DEPRECATED __attribute__((__deprecated__))
#define OGRE_DEPRECATED_EXPORT _OgreExport OGRE_DEPRECATED
#define OGRE_DEPRECATED_NO_EXPORT _OgrePrivate OGRE_DEPRECATED
#endif
