md5 : b660f56779429a5f4bdf380eb7025073
name : txt2vcf_pl_b660
version : ""
language : C
language_id : C
installed : 0
date_compiled : Tue Oct 20 04:37:49 2026
inline_version : 0.53
ILSM : %
    module : Inline::C
    suffix : so
    type : compiled
Config : %
    apiversion : ?
    archname : x86_64-linux-gnu-thread-multi
    cc : x86_64-linux-gnu-gcc
    ccflags : -D_REENTRANT -D_GNU_SOURCE -DDEBIAN -fwrapv -fno-strict-aliasing -pipe -I/usr/local/include -D_LARGEFILE_SOURCE -D_FILE_OFFSET_BITS=64
    ld : x86_64-linux-gnu-gcc
    osname : linux
    osvers : 6.1.0
    so : so
    version : 5.38.2
