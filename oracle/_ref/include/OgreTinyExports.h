#ifndef OGRE_TINY_EXPORTS_H
#define OGRE_TINY_EXPORTS_H
#define _OgreTinyExport __attribute__((visibility("default")))
#endif
