/* force-included (-include) when glue/Makefile compiles the reference's src_main/xevdm_alf.c with the definition of
 * alf_process_tile renamed: alf_process (xevdm_alf.c:1167) then binds to the device-launching alf_process_tile of
 * glue/xevd_b200_glue.c */
int alf_process_tile(void *arg);
