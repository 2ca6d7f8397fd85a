#pragma once
typedef struct CUtensorMap_st { alignas(64) unsigned long long opaque[16]; } CUtensorMap;
