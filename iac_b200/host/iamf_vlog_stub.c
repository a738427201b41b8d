/*
 * iamf_vlog_stub.c - the "verification log" entry points the stock iamfplayer links against
 * (Samsung/iac include/vlogging_tool_sr.h:46-58; the generator itself, src/iamf_dec/vlogging_tool_sr.c, is conformance
 * tooling outside the rendering path).  The reference library exports them unconditionally and the player's MP4 code
 * references them even when SUPPORT_VERIFIER is off, so a drop-in libiamf.so must resolve them.  Here logging is
 * permanently "not open": opening fails, nothing is ever written.
 */
#include <stdarg.h>
#include <stdint.h>

typedef enum LOG_TYPE { LOG_OBU = 0, LOG_MP4BOX = 1, LOG_DECOP = 2, MAX_LOG_TYPE } LOG_TYPE;

int vlog_file_open(const char *log_file_name) { (void)log_file_name; return -1; }
int vlog_file_close(void) { return 0; }
int is_vlog_file_open(void) { return 0; }
int vlog_print(LOG_TYPE type, uint64_t key, const char *format, ...) { (void)type; (void)key; (void)format; return 0; }
int vlog_obu(uint32_t obu_type, void *obu, uint64_t trim_start, uint64_t trim_end) {
  (void)obu_type; (void)obu; (void)trim_start; (void)trim_end;
  return 0;
}
int write_prefix(LOG_TYPE type, char *buf) { (void)type; if (buf) buf[0] = 0; return 0; }
int write_postfix(LOG_TYPE type, char *buf) { (void)type; (void)buf; return 0; }
int write_yaml_form(char *log, uint8_t indent, const char *format, ...) { (void)indent; (void)format; if (log) log[0] = 0; return 0; }
